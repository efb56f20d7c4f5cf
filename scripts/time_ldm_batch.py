"""Large-data mode, the normalise + PCA of the partitions of one rank: all partitions queued together over CUDA
streams (PartitionPCA, what cluster_by_taxon_partitioning does) against one partition after the other
(a) on one stream on the device and (b) through the reference-shaped calls kmers.normalize + kmers.pca.
Synthetic: 200k contigs, k = 5 counts, 100 partitions of 2000 contigs (max_partition_size = 10000)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from autometa_b200 import assembly, device as D, synth
from autometa_b200.binning.large_data_mode import PartitionPCA
from autometa_b200.common import kmers

n_parts, per = 100, 2000
asm = synth.make_assembly(n_parts * per, n_parts * per * 4000, seed=9)
packed = assembly.pack_to_device(asm, device="cuda:0", on_device=True)
cnt = D.count_packed(packed, 5)[0].cpu().numpy().astype(np.float64)
cnt[cnt == 0] = np.nan
counts = pd.DataFrame(cnt, index=pd.Index(asm.ids, name="contig"), columns=D.column_names(5))
parts = {f"t{i}": counts.index[i * per:(i + 1) * per] for i in range(n_parts)}
out = {"partitions": n_parts, "contigs_per_partition": per, "k": 5, "norm": "am_clr", "pca_dimensions": 50}
for streams in (8, 1):
    pp = PartitionPCA(counts, "am_clr", 50, n_streams=streams)
    pp.run(dict(list(parts.items())[:10]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = pp.run(parts)
    torch.cuda.synchronize()
    out[f"partition_pca_{streams}_streams_s"] = round(time.perf_counter() - t0, 4)
t0 = time.perf_counter()
for name, idx in list(parts.items())[:20]:
    norm = kmers.normalize(counts.loc[idx], method="am_clr")
    kmers.pca(norm.to_numpy(), n_components=50)
torch.cuda.synchronize()
out["normalize_plus_pca_calls_s_scaled_to_100"] = round((time.perf_counter() - t0) * n_parts / 20, 4)
from sklearn.decomposition import PCA
t0 = time.perf_counter()
for name, idx in list(parts.items())[:5]:
    X = kmers.normalize(counts.loc[idx], method="am_clr").to_numpy()
    PCA(n_components=50, random_state=np.random.RandomState(42)).fit_transform(X)
out["gpu_normalize_plus_sklearn_pca_s_scaled_to_100"] = round((time.perf_counter() - t0) * n_parts / 5, 4)
print(json.dumps(out))
