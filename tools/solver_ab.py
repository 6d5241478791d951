import os, sys, json
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "quantum-targeted-energy-transfer_b200"))
import torch
import bench
from tet import _native
dev = torch.device("cuda", 0)
for rep in range(2):
    r = bench.run_solver_block(_native, torch, dev, 0, 0, 1)
    print(os.environ.get("QTET_ADAM_RESORT", "default"), "wall %.4f evals %d best %r idx %d launches %d" % (r["wall_s"], r["loss_grad_evals"], r["best_loss"], r["best_index"], _native.launch_count()))
