"""bf16 / fp16 logit error of the CUDA path vs the CPU oracle for a 15-frame clip (ED and VAE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from model.config import load_config
from model.genconvit_ed import GenConViTED
from model.genconvit_vae import GenConViTVAE
from oracle import nets
from oracle.weights import make_state_dict, synthetic_eps, synthetic_frames
cfg = load_config(); dev = "cuda"
n = int(os.environ.get("N", "15"))
x, eps = synthetic_frames(n, 21), synthetic_eps(n, 22)
with torch.no_grad():
    sd = make_state_dict("ed", 0)
    want = nets.ed_forward(sd, x)
    m = GenConViTED(cfg).eval(); m.load_state_dict(sd); m.to(dev)
    for mode in ("bf16", "fp16"):
        got = m.set_compute_dtype(mode)(x.to(dev)).float().cpu()
        print(f"ED  {mode}: max|dlogit| = {(got-want).abs().max():.3e}  mean = {(got-want).abs().mean():.3e}  margin_min = {(want[:,0]-want[:,1]).abs().min():.3f}")
    del m, sd
    sd = make_state_dict("vae", 0, skip_var=True)
    want = nets.vae_forward(sd, x, eps, resize=False)[0]
    m = GenConViTVAE(cfg).eval(); m.load_state_dict(sd); m.to(dev)
    for mode in ("bf16", "fp16"):
        got = m.set_compute_dtype(mode)(x.to(dev), eps=eps.to(dev))[0].float().cpu()
        print(f"VAE {mode}: max|dlogit| = {(got-want).abs().max():.3e}  mean = {(got-want).abs().mean():.3e}")
