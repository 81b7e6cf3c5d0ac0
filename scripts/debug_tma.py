"""Debug helper: run one warp_fuse configuration in this process and compare with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gencomm_b200 as G
from gencomm_b200 import synth
from oracle import ref_ops as R

C, H, W, N, mode = [int(v) for v in sys.argv[1:6]]
feat = synth.bev_features(1, N, C, H, W)
pw = synth.pairwise_t_matrix(1, N, 5, spread=(0.3 * W * 0.4, 0.3 * H * 0.4))[None]
theta = R.normalize_pairwise_tfm(torch.from_numpy(pw), H * 0.4, W * 0.4, 1)
rl = torch.tensor([N])
fd, rd, td = feat.cuda(), rl.cuda(), theta.cuda()
fn = [G.warp_feature, G.MaxFusion(), G.AttFusion(C)][mode]
ref = [R.warp_only, R.max_fusion, R.att_fusion][mode](feat, rl, theta)
out = fn(fd, rd, td)
torch.cuda.synchronize()
err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
print(f"C={C} H={H} W={W} N={N} mode={mode} GATHER={os.environ.get('GC_WARP_FUSE_GATHER')} rel_err={err:.3e}")
