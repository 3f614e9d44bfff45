"""Profiling driver: k-means(100) fits on one 4096 x 2048 synthetic slide (run under ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kmeans_oracle as K
from sequoia_pub_b200.kmeans import KMeans
X = torch.from_numpy(K.make_slide_features(0)).cuda()
km = KMeans(n_clusters=100, random_state=0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    km.fit(X)
torch.cuda.synchronize()
