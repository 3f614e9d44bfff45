#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_kmeans_gpu.py -m gpu -x -q -k "odd or edge" > gpurun_out/sanitizer_kmeans.log 2>&1; tail -5 gpurun_out/sanitizer_kmeans.log
python -m pytest tests/test_kmeans_gpu.py -m gpu -q > gpurun_out/pytest_kmeans.log 2>&1; tail -30 gpurun_out/pytest_kmeans.log
python - <<'PY' 2>&1 | tee gpurun_out/kmeans_timing.log
import sys, time, torch
sys.path.insert(0, '.')
from oracle import kmeans_oracle as K
from sequoia_pub_b200.kmeans import KMeans
X = torch.from_numpy(K.make_slide_features(0)).cuda()
km = KMeans(n_clusters=100, random_state=0)
km.fit(X); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): km.fit(X)
torch.cuda.synchronize()
print(f"[kmeans] {(time.perf_counter()-t0)/5*1e3:.2f} ms per slide (4096x2048, k=100, {km.n_iter_} Lloyd iterations)")
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kmeans.csv python - <<'PY' > /dev/null 2>&1
import sys, torch
sys.path.insert(0, '.')
from oracle import kmeans_oracle as K
from sequoia_pub_b200.kmeans import KMeans
X = torch.from_numpy(K.make_slide_features(0)).cuda()
KMeans(n_clusters=100, random_state=0).fit(X)
PY
