"""The reference's own unit tests for this path, against the drop-in (tests/unit_tests/test_kmers.py:77-223 and
test_recursive_dbscan.py:113-165 of the reference, same names and assertions).  Fixtures the reference takes from
its (absent) tests/data are replaced by the committed small FASTA and the seeded synthetic binning table."""
import sys
import types

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kmers():
    import torch

    assert torch.cuda.is_available()
    from autometa_b200.common import kmers as mod

    return mod


@pytest.fixture(scope="module")
def assembly(small_fna):
    return small_fna


@pytest.fixture(scope="module")
def counts(kmers, assembly):
    return kmers.count(assembly=assembly, size=5)


@pytest.fixture(scope="module")
def counts_fpath(counts, tmp_path_factory):
    p = tmp_path_factory.mktemp("kmers") / "counts.tsv"
    counts.to_csv(p, sep="\t", index=True, header=True)
    return str(p)


@pytest.fixture(scope="module")
def norm_df(kmers, counts):
    return kmers.normalize(df=counts, method="am_clr")


@pytest.fixture(scope="module")
def invalid_df_fpath(tmp_path_factory):
    p = tmp_path_factory.mktemp("kmers") / "invalid_df.tsv"
    pd.DataFrame({"column1": ["invalid_contig_1", "invalid_contig_2"], "column2": ["invalid_marker1", "invalid_marker2"]}
                 ).to_csv(p, sep="\t")
    return str(p)


@pytest.fixture(autouse=True)
def bh_sne_stub(monkeypatch):
    # the `tsne` package is not installed: embed's manifold step is not on the path under test
    stub = types.ModuleType("tsne")
    stub.bh_sne = lambda data, d, **kw: np.ascontiguousarray(data[:, :d])
    monkeypatch.setitem(sys.modules, "tsne", stub)


def test_kmer_load(kmers, counts_fpath):
    df = kmers.load(kmers_fpath=counts_fpath)
    assert not df.empty
    assert df.index.name == "contig"


def test_kmer_load_missing_file(kmers):
    with pytest.raises(FileNotFoundError):
        kmers.load(kmers_fpath="Invalid_fpath")


def test_kmer_load_table_missing_contig_column(kmers, invalid_df_fpath):
    from autometa_b200.common.exceptions import TableFormatError

    with pytest.raises(TableFormatError):
        kmers.load(invalid_df_fpath)


def test_count(kmers, assembly, tmp_path):
    out = tmp_path / "kmers.tsv"
    size = 5
    df = kmers.count(assembly=assembly, size=size, out=out, force=False)
    assert df.shape[1] == 4 ** size / 2
    assert df.index.name == "contig"
    assert out.exists()


@pytest.mark.parametrize("force", [True, False])
def test_count_out_exists(kmers, assembly, counts, force, tmp_path):
    out = tmp_path / "kmers.tsv"
    counts.to_csv(out, sep="\t", index=True, header=True)
    df = kmers.count(assembly=assembly, size=5, out=out, force=force)
    assert df.shape[1] == 4 ** 5 / 2
    assert df.index.name == "contig"
    assert out.exists()


def test_count_wrong_size(kmers, assembly):
    with pytest.raises(TypeError):
        kmers.count(assembly=assembly, size=5.5)


@pytest.mark.parametrize("method", ["am_clr", "clr", "ilr"])
def test_normalize(kmers, counts, method, tmp_path):
    out = tmp_path / "kmers.norm.tsv"
    table = counts if method == "am_clr" else counts.dropna(axis="index", how="all")  # skbio rejects all-zero rows
    df = kmers.normalize(df=table, method=method, out=out, force=False)
    if method == "am_clr":
        assert df.shape[1] == counts.shape[1]
    elif method == "clr":
        assert df.shape == table.shape
    else:
        assert df.shape[1] < counts.shape[1]  # ILR will reduce the columns by one.
    assert out.exists()


@pytest.mark.parametrize("force", [True, False])
def test_normalize_out_exists(kmers, counts, norm_df, force, tmp_path):
    out = tmp_path / "kmers.norm.tsv"
    norm_df.to_csv(out, sep="\t", index=True, header=True)
    df = kmers.normalize(df=counts, method="am_clr", out=out, force=force)
    assert df.shape == norm_df.shape
    assert df.index.name == "contig"


def test_normalize_wrong_method(kmers, counts, tmp_path):
    with pytest.raises(ValueError):
        kmers.normalize(df=counts, method="am_ilr", out=tmp_path / "kmers.norm.tsv", force=False)


@pytest.mark.parametrize("method", ["bhsne", "sksne", "umap", "densmap"])
def test_embed_methods(kmers, norm_df, method, tmp_path):
    if method in ("umap", "densmap"):
        pytest.importorskip("umap")
    method_kwargs = {"verbose": 1 if method == "sksne" else True} if method != "bhsne" else {}
    if method == "densmap":
        method_kwargs["output_dens"] = True
    df = kmers.embed(kmers=norm_df, out=tmp_path / "kmers.embed.tsv", force=False, embed_dimensions=2,
                     pca_dimensions=3, method=method, seed=42, **method_kwargs)
    assert df.shape[1] == (4 if method == "densmap" else 2)


@pytest.mark.parametrize("embed_dimensions", [2, 3, 4])
def test_embed_dimensions(kmers, norm_df, embed_dimensions, tmp_path):
    df = kmers.embed(kmers=norm_df, out=tmp_path / "kmers.embed.tsv", force=False, embed_dimensions=embed_dimensions,
                     pca_dimensions=5, method="bhsne", seed=42)
    assert df.shape[1] == embed_dimensions


@pytest.mark.parametrize("force", [False, True])
def test_embed_out_exists(kmers, norm_df, force, tmp_path):
    out = tmp_path / "kmers.embed.tsv"
    first = kmers.embed(kmers=norm_df, out=out, force=True, embed_dimensions=2, pca_dimensions=3, method="bhsne", seed=42)
    df = kmers.embed(kmers=norm_df, out=out, force=force, embed_dimensions=2, pca_dimensions=3, method="bhsne", seed=42)
    assert df.shape == first.shape


def test_embed_TableFormatError(kmers, invalid_df_fpath):
    from autometa_b200.common.exceptions import TableFormatError

    with pytest.raises(TableFormatError):
        kmers.embed(kmers=invalid_df_fpath)


def test_embed_input_not_string_or_dataframe(kmers, tmp_path):
    with pytest.raises(TypeError):
        kmers.embed(kmers=tmp_path / "kmers.embed.tsv")


def test_embed_empty_dataframe(kmers, tmp_path):
    with pytest.raises(FileNotFoundError):
        kmers.embed(kmers=pd.DataFrame({}), out=tmp_path / "kmers.embed.tsv", force=True)


# ---- recursive_dbscan (reference tests/unit_tests/test_recursive_dbscan.py:113-165)
@pytest.fixture(scope="module")
def binning_inputs():
    from autometa_b200 import synth

    table, markers = synth.make_sweep_input(n_points=1500, n_clusters=9, n_markers=40, seed=21)
    return synth.add_taxonomy(table, seed=3), markers


@pytest.mark.parametrize("method", ["dbscan"])
def test_taxon_guided_binning(binning_inputs, method):
    from autometa_b200.binning import recursive_dbscan

    main_df, markers_df = binning_inputs
    df = recursive_dbscan.taxon_guided_binning(
        main=main_df.copy(), markers=markers_df, completeness=20.0, purity=95.0, coverage_stddev=25.0,
        gc_content_stddev=5.0, method=method)
    for col in ["cluster", "purity", "completeness", "coverage_stddev", "gc_content_stddev"]:
        assert col in df.columns
    assert df.shape[0] == main_df.shape[0]


def test_taxon_guided_binning_invalid_clustering_method(binning_inputs):
    from autometa_b200.binning import recursive_dbscan

    main_df, markers_df = binning_inputs
    with pytest.raises(ValueError):
        recursive_dbscan.taxon_guided_binning(main=main_df.copy(), markers=markers_df, method="invalid_clustering_method")


def test_taxon_guided_binning_empty_markers_table(binning_inputs):
    from autometa_b200.binning import recursive_dbscan
    from autometa_b200.common.exceptions import TableFormatError

    main_df, markers_df = binning_inputs
    with pytest.raises(TableFormatError):
        recursive_dbscan.taxon_guided_binning(main=main_df.copy(), markers=markers_df.iloc[:0], method="dbscan")
