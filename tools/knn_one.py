"""One K=16 self-query on a 180k uniform cloud (for ncu captures of knn_search_kernel).  PU_KNN_DEFER=1 selects the queued insertion."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from point_unet_b200.helper_tool import knn_search_cuda, knn_last_stats
n = int(sys.argv[1]) if len(sys.argv) > 1 else 180000
x = torch.from_numpy(np.random.default_rng(n).random((1, n, 3), dtype=np.float32)).cuda()
for _ in range(2):
    out = knn_search_cuda(x, x, 16)
torch.cuda.synchronize()
print(knn_last_stats(), int(out.sum()))
