"""CPU checks of the cluster-kernel weight records (gencomm.pack_unet_cluster): a numpy emulation of the
input-row-stationary tensor-core formulation of csrc/denoiser_cluster.cu (one [128 x cin] x [cin x 32] product per
staged input row and tap kx, block j of the 32 columns accumulating into output row i - 2 + j) must reproduce
torch's 3x3 convolution with the reference weights (unet.py:117-138)."""
import numpy as np
import torch

import gencomm_b200 as G
from gencomm_b200 import gencomm as GM


def _cfg(C):
    return {"model": {"embed_dim": C + 2, "in_channels": C, "out_ch": C, "ch": 8, "ch_mult": [1, 1],
                      "num_res_blocks": 2, "attn_resolutions": [16], "dropout": 0.0, "resamp_with_conv": True},
            "diffusion": {"beta_schedule": "linear", "beta_start": 0.0005, "beta_end": 0.02,
                          "num_diffusion_timesteps": 3}}


def _emulate(rec, x, cin):
    """x [cin, R, 128] (rows outside are zero padding) -> [8, R, 128] through the record's B operand."""
    R = x.shape[1]
    b = rec[:1536].reshape(3, 2, 2, 4, 8, 4)                       # [kx][cg][half][j][cout][cin4]
    rows = np.zeros((R + 2, 130, cin), dtype=np.float64)            # staged rows i = 0..R+1 <-> y = i - 1, px = x + 1
    rows[1:R + 1, 1:129, :] = x.transpose(1, 2, 0)
    acc = np.zeros((128, 8 * (R + 2 + 3)), dtype=np.float64)        # TMEM: lane = pixel, column 8 (r + 2) + cout
    for i in range(R + 2):
        for kx in range(3):
            a = rows[i, kx:kx + 128, :]                             # [128, cin]
            bm = b[kx, :cin // 8].transpose(2, 3, 0, 1, 4).reshape(32, cin)   # [j*8+cout][cg*8+half*4+e]
            acc[:, 8 * i:8 * i + 32] += a @ bm.T
    out = np.stack([acc[:, 8 * (r + 2):8 * (r + 2) + 8] for r in range(R)])   # [R, 128, 8]
    return out.transpose(2, 0, 1)


def test_cluster_records_reproduce_the_convolutions():
    torch.manual_seed(3)
    m = G.GenComm(_cfg(16))
    sd = {k[len("denoiser."):]: v.detach() for k, v in m.state_dict().items() if k.startswith("denoiser.")}
    host, _ = GM.pack_unet(sd, 16, 3)
    cl = GM.pack_unet_cluster(host, 3).reshape(3, GM.N_C8_LAYERS, GM.CL_REC_FLOATS)
    names = []
    for name, kind in GM._LAYERS:
        names += [(name, "plain")] if kind == "plain" else [(name + ".conv1", kind), (name + ".conv2", "res8")]
    rng = np.random.default_rng(0)
    for li, (name, kind) in enumerate(names):
        w = sd[name + ".weight"] if kind != "plain" else sd[name + ".weight"]
        cin = w.shape[1]
        rec = cl[1, li]
        if li == GM._DOWN_LAYER:
            ref = w.permute(2, 3, 1, 0).reshape(9, 8, 8).numpy()   # [tap][cin][cout]
            assert np.array_equal(rec[:576].reshape(9, 8, 8), ref)
            continue
        x = rng.standard_normal((cin, 5, 128)).astype(np.float32)
        ref = torch.nn.functional.conv2d(torch.from_numpy(x)[None].double(), w.double(), padding=1)[0].numpy()
        out = _emulate(rec.astype(np.float64), x.astype(np.float64), cin)
        assert np.abs(out - ref).max() <= 1e-10, (li, name)
        assert np.all(rec[:1536].reshape(3, 2, 2, 4, 8, 4)[:, :, :, 3] == 0)   # the fourth column block is zero
    # bias carries the timestep-embedding projection of its step; GroupNorm affine / nin_shortcut are copied through
    for t in range(3):
        rec = cl[t, 0]
        assert not np.array_equal(cl[0, 0][1536:1544], cl[2, 0][1536:1544])
        assert np.array_equal(rec[1544:1552], sd["down.0.block.0.norm1.weight"].numpy())
    nin = sd["up.1.block.0.nin_shortcut.weight"].reshape(8, 16).numpy()
    assert np.array_equal(cl[0, 14][1576:1704].reshape(16, 8), nin.T)
