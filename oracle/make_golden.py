"""Oracle (test infrastructure): generate tests/golden/*.pt from the REAL reference.

Runs only in the authoring container, where ``/root/reference`` exists:

    python -m oracle.make_golden            # writes tests/golden/{ed,vae,genconvit}_golden.pt

It imports the reference's own, unmodified ``model/genconvit*.py`` and
``model/pred_func.py`` (through symlinks in a scratch directory, nothing is
copied into the repo), with ``oracle.timm_standin`` standing in for the absent
timm==0.6.5 and stubs for dlib / face_recognition / decord / albumentations.
Weights are the seeded state_dicts of ``oracle.weights`` written as
``weight/*.pth`` (ED raw, VAE wrapped in ``{'state_dict': ...}`` to exercise both
load paths of reference model/genconvit.py:18-21) and loaded with
``strict=True`` -- which also proves the key inventory in ``oracle.weights``.

The VAE's ``torch.randn_like`` (reference model/genconvit_vae.py:46) is patched to
return the injected epsilon so the stochastic reference becomes reproducible.

The goldens hold only seeds + small outputs (logits, scores, a few strided
intermediate samples); tests regenerate inputs/weights from the same seeds.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _scratch():
    d = tempfile.mkdtemp(prefix="gcv_ref_")
    os.symlink(os.path.join(REF, "model"), os.path.join(d, "model"))
    os.symlink(os.path.join(REF, "dataset"), os.path.join(d, "dataset"))
    os.makedirs(os.path.join(d, "weight"))
    return d


def main(n_ed=3, n_vae=2, n_vid=15, seed=0):
    from oracle import nets, timm_standin
    from oracle.weights import make_state_dict, synthetic_eps, synthetic_frames

    torch.set_grad_enabled(False)
    scratch = _scratch()
    os.chdir(scratch)
    sys.path.insert(0, scratch)
    timm_standin.install()

    sd_ed = make_state_dict("ed", seed)
    sd_vae = make_state_dict("vae", seed)
    torch.save(sd_ed, "weight/ed_rand.pth")
    torch.save({"state_dict": sd_vae}, "weight/vae_rand.pth")

    import model.genconvit_vae as ref_vae_mod           # the reference's files, via the symlink
    from model.config import load_config
    from model.genconvit import GenConViT
    from model.genconvit_ed import GenConViTED
    from model.genconvit_vae import GenConViTVAE
    import model.pred_func as ref_pred
    assert ref_vae_mod.__file__.startswith(scratch)

    config = load_config()
    os.makedirs(OUT, exist_ok=True)
    meta = dict(seed=seed, torch=str(torch.__version__), generated=time.strftime("%Y-%m-%d"),
                reference_files=["model/genconvit.py", "model/genconvit_ed.py", "model/genconvit_vae.py",
                                 "model/model_embedder.py", "model/pred_func.py"])

    # ---- ED ----------------------------------------------------------------
    ed = GenConViTED(config, pretrained=False).eval()
    print("ED state_dict entries:", len(ed.state_dict()), ed.load_state_dict(sd_ed, strict=True))
    x = synthetic_frames(n_ed, seed)
    ref_logits = ed(x)
    taps = {}
    ora_logits = nets.ed_forward(sd_ed, x, taps)
    print("ED  oracle vs reference max|d| =", (ref_logits - ora_logits).abs().max().item())
    emb = ed.embedder(x[:1])
    torch.save(dict(meta=meta, n=n_ed, logits=ref_logits, n_entries=len(ed.state_dict()),
                    decoded_sample=taps["decoded"][:, :, ::37, ::41].clone(),
                    x1_sample=taps["x1"][:, ::50].clone(), x2_sample=taps["x2"][:, ::50].clone(),
                    embedder_logits_sample=emb[:, ::50].clone(),
                    oracle_vs_reference=(ref_logits - ora_logits).abs().max().item()),
               os.path.join(OUT, "ed_golden.pt"))

    # ---- VAE ---------------------------------------------------------------
    vae = GenConViTVAE(config, pretrained=False).eval()
    print("VAE state_dict entries:", len(vae.state_dict()), vae.load_state_dict(sd_vae, strict=True))
    x = synthetic_frames(n_vae, seed + 1)
    eps = synthetic_eps(n_vae, 7)
    real_randn_like = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps.to(t.dtype)
    try:
        ref_logits, ref_xhat = vae(x)
        ref_kl = vae.encoder.kl.clone()
    finally:
        torch.randn_like = real_randn_like
    taps = {}
    ora_logits, ora_xhat = nets.vae_forward(sd_vae, x, eps, taps)
    ora_kl = nets.vae_kl(sd_vae, taps["feat"], taps["mu"])
    print("VAE oracle vs reference max|d| logits =", (ref_logits - ora_logits).abs().max().item(),
          " x_hat =", (ref_xhat - ora_xhat).abs().max().item(), " kl =", (ref_kl - ora_kl).abs().item())
    torch.save(dict(meta=meta, n=n_vae, eps_seed=7, frames_seed=seed + 1, logits=ref_logits, kl=ref_kl,
                    n_entries=len(vae.state_dict()),
                    xhat224_sample=ref_xhat[:, :, ::37, ::41].clone(),
                    mu_sample=taps["mu"][:, ::1001].clone(), z_sample=taps["z"][:, ::1001].clone(),
                    x1_sample=taps["x1"][:, ::50].clone(), x2_sample=taps["x2"][:, ::50].clone(),
                    oracle_vs_reference=(ref_logits - ora_logits).abs().max().item()),
               os.path.join(OUT, "vae_golden.pt"))
    del vae, ed

    # ---- full GenConViT through the reference's pred_vid --------------------
    model = GenConViT(config, ed="ed_rand", vae="vae_rand", net="genconvit", fp16=False).eval()
    x = synthetic_frames(n_vid, seed + 2)
    eps = synthetic_eps(n_vid, 8)
    torch.randn_like = lambda t, *a, **k: eps.to(t.dtype)
    try:
        ref_rows = model(x)
        ref_score = ref_pred.pred_vid(x, model)
    finally:
        torch.randn_like = real_randn_like
    ora_rows = nets.genconvit_forward(sd_ed, sd_vae, x, eps)
    ora_score = nets.pred_vid(ora_rows)
    print("full oracle vs reference max|d| =", (ref_rows - ora_rows).abs().max().item(), ref_score, ora_score)
    torch.save(dict(meta=meta, n=n_vid, frames_seed=seed + 2, eps_seed=8, rows=ref_rows,
                    pred_vid=(int(ref_score[0]), float(ref_score[1])),
                    real_or_fake=ref_pred.real_or_fake(ref_score[0]),
                    oracle_vs_reference=(ref_rows - ora_rows).abs().max().item()),
               os.path.join(OUT, "genconvit_golden.pt"))
    print("margins |l0-l1| min/median:", (ref_rows[:, 0] - ref_rows[:, 1]).abs().min().item(),
          (ref_rows[:, 0] - ref_rows[:, 1]).abs().median().item(), " logits std", ref_rows.std().item())
    print("wrote", OUT)


if __name__ == "__main__":
    main()
