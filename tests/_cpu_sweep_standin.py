"""CPU stand-in for ``autometa_b200.device.DeviceSweep`` (tests only): the reference's sweep loop
(recursive_dbscan.py:172-215) over the numpy oracle, behind DeviceSweep's interface, so that the host logic built
on top of the device sweep — result frames, ``get_clusters`` rounds, bin naming, the in-flight driver,
``taxon_guided_binning`` — is checked against the reference's golden frames without a GPU."""
import numpy as np

from oracle import oracle_np as O


class OracleSweep:
    def __init__(self, X, coverage, gc, marker_rows, marker_cols, marker_vals, n_ref_markers, cutoffs, device,
                 stream=None, trace_cap=0, graph=True):
        n = X.shape[0]
        markers = np.zeros((n, int(n_ref_markers)), dtype=np.int64)
        if len(marker_rows):
            np.add.at(markers, (np.asarray(marker_rows), np.asarray(marker_cols)), np.asarray(marker_vals).astype(np.int64))
        cc, pc, cs, gs = cutoffs
        eps, step = 0.3, 0.1
        clustering_rounds = {0: 0}
        n_clusters = float("inf")
        best_median = float("-inf")
        best = None
        self.steps = 0
        while n_clusters > 1:
            labels = O.dbscan_labels(X, eps)
            m = O.cluster_metrics(labels, markers, coverage, gc, int(n_ref_markers))
            with np.errstate(invalid="ignore"):
                ok = ((m["completeness"] >= cc) & (m["purity"] >= pc) & (m["coverage_stddev"] <= cs)
                      & (m["gc_content_stddev"] <= gs))
            median = float(np.median(m["completeness"][ok])) if ok.any() else float("-inf")
            if median >= best_median:
                best_median = median
                best = (labels, m)
            n_clusters = int(labels.max()) + 1
            clustering_rounds[n_clusters] = clustering_rounds.get(n_clusters, 0) + 1
            if clustering_rounds[n_clusters] > 10:
                step *= 10
            if median == 0:
                clustering_rounds[0] += 1
            self.steps += 1
            if clustering_rounds[0] >= 10:
                break
            eps += step
        self._result = (best[0].astype(np.int64),
                        {k: (v.astype(np.int64) if k.endswith("count") else v) for k, v in best[1].items()},
                        best_median, int(best[0].max()) + 1, None)
        self._launched = 0

    # the part of DeviceSweep's interface the drivers use
    def launch(self, steps=0):
        self._launched += steps or 16

    def ready(self):
        return True

    def wait(self):
        pass

    def poll(self):
        return self._launched >= self.steps

    def result(self):
        return self._result

    def close(self):
        pass


def install(monkeypatch):
    """Route recursive_dbscan's device calls to the stand-in (no CUDA needed)."""
    import torch

    from autometa_b200 import device as D
    from autometa_b200.binning import recursive_dbscan as R

    monkeypatch.setattr(D, "DeviceSweep", OracleSweep)
    monkeypatch.setattr(R.L, "require_cuda", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(R, "_stream_pool", lambda dev, k: [None] * k)


class OraclePartitionPCA:
    """CPU stand-in for ``large_data_mode.PartitionPCA`` (tests only): am_clr + exact PCA of the numpy oracle per
    partition, rows in the count table's order."""

    def __init__(self, counts, norm_method, pca_dimensions, device=None, n_streams=8):
        assert norm_method == "am_clr"
        self.counts, self.P = counts, int(pca_dimensions)

    def run(self, partitions):
        import pandas as pd

        out = {}
        for key, index in partitions.items():
            sub = self.counts.loc[self.counts.index.isin(index)]
            X, row_keep, _ = O.am_clr(sub.to_numpy(dtype=np.float64))
            out[key] = pd.DataFrame(O.pca_cov_eigh(X, self.P)["scores"], index=sub.index[row_keep])
        return out


def install_ldm(monkeypatch):
    """``install`` + the batched normalise / PCA of large-data mode and the ``tsne`` stub the reference's golden
    run used (returns the first d columns it is handed)."""
    import sys
    import types

    from autometa_b200.binning import large_data_mode as M

    install(monkeypatch)
    monkeypatch.setattr(M, "PartitionPCA", OraclePartitionPCA)
    stub = types.ModuleType("tsne")
    stub.bh_sne = lambda data, d, **kw: np.ascontiguousarray(data[:, :d])
    monkeypatch.setitem(sys.modules, "tsne", stub)
