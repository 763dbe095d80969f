"""Scoring / loading helpers -- drop-in for reference model/pred_func.py.

Same public names and return conventions (``load_genconvit``, ``preprocess_frame``,
``pred_vid``, ``max_prediction_value``, ``real_or_fake``, ``set_result``,
``store_result``, ``is_video``, ``df_face``, ``device``).  Differences, all deliberate:

* importing this module does not need dlib / face_recognition / decord / albumentations /
  timm: face extraction and video decoding (reference :67-92, :138-149) stay host-side and are
  imported lazily by ``df_face`` only;
* ``pred_vid`` scores on the GPU with one fused kernel and ONE device->host copy instead of
  three ``.item()`` syncs (reference :123-131), and works for N = 1 (the reference's
  ``.squeeze()`` breaks there);
* ``pred_videos`` scores many videos from one batched forward (the reference runs one
  forward per video).
"""
import os

import numpy as np
import torch

from genconvit_b200 import engine

from .config import load_config  # noqa: F401  (re-exported like the reference)
from .genconvit import GenConViT

device = "cuda" if torch.cuda.is_available() else "cpu"

_MEAN = (0.485, 0.456, 0.406)     # dataset/loader.py:64-65 (normalize_data()["vid"])
_STD = (0.229, 0.224, 0.225)


def load_genconvit(config, net, ed_weight, vae_weight, fp16, arch_type="original", use_attention=True,
                   use_residual=True):
    """reference :18-64.  Only the original architecture exists here (GenConViTV2 is an unused
    wrapper duplicate in the reference, out of scope)."""
    if arch_type != "original":
        raise NotImplementedError("arch_type='v2' (GenConViTV2) is out of scope; use arch_type='original'")
    model = GenConViT(config, ed=ed_weight, vae=vae_weight, net=net, fp16=fp16)
    model.to(device)
    model.eval()
    if fp16:
        model.half()
    return model


def preprocess_frame(frame):
    """uint8 NHWC face crops -> ImageNet-normalised fp32 NCHW on ``device`` (reference :95-108).

    With a GPU the uint8 frames are copied as they are (a quarter of the fp32 bytes) and normalised by
    the ``gcv_preprocess_frames`` kernel -- same fp32 operation order as the reference's per-frame CPU loop, so the
    result is bit-identical; without one this is host pre-processing, done with torch like the reference."""
    arr = np.ascontiguousarray(np.asarray(frame))
    if torch.cuda.is_available() and arr.dtype == np.uint8 and arr.ndim == 4 and arr.shape[3] == 3 \
            and (arr.shape[1] * arr.shape[2]) % 4 == 0 and arr.shape[0] > 0:
        from genconvit_b200 import lib as L
        n, h, w, _ = arr.shape
        u8 = torch.from_numpy(arr).to(device, non_blocking=False)
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=device)
        L.preprocess_frames(u8, out, n, h, w, _MEAN, _STD)
        return out
    df = torch.as_tensor(arr).float().permute(0, 3, 1, 2) / 255.0
    mean = torch.tensor(_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(_STD).view(1, 3, 1, 1)
    df = (df - mean) / std
    return df.to(device) if torch.cuda.is_available() else df


def _model_device(model):
    return next(model.parameters()).device


def pred_vid(df, model):
    """One video: forward + sigmoid + mean over rows + argmax, returns ``(class, score)``
    exactly as reference :111-131 does (class 0 = FAKE, 1 = REAL after ``real_or_fake``)."""
    with torch.no_grad():
        dev = _model_device(model)
        if df.device != dev:
            df = df.to(dev)
        logits = model(df).float().reshape(-1, 2).contiguous()
        n_nets = 2 if getattr(model, "net", None) not in ("ed", "vae") and logits.shape[0] == 2 * df.shape[0] else 1
        n = logits.shape[0] // n_nets
        _, cls, val = engine.score_videos(logits, n_nets, n, n)
        out = torch.stack((cls.float(), val)).cpu()            # the single device->host copy
        return int(out[0, 0].item()), float(out[1, 0].item())


def pred_videos(df, model, frames_per_video):
    """Batched form of ``pred_vid``: ``df`` holds V videos of ``frames_per_video`` consecutive frames.
    Returns (classes [V] int tensor, scores [V] float tensor) on the host."""
    with torch.no_grad():
        dev = _model_device(model)
        if df.device != dev:
            df = df.to(dev)
        logits = model(df).float().reshape(-1, 2).contiguous()
        n_nets = logits.shape[0] // df.shape[0]
        _, cls, val = engine.score_videos(logits, n_nets, df.shape[0], frames_per_video)
        out = torch.stack((cls.float(), val)).cpu()
        return out[0].to(torch.int64), out[1]


def max_prediction_value(y_pred):
    """reference :123-131 on an already-sigmoided [rows,2] tensor (host-side utility form)."""
    mean_val = torch.mean(y_pred.float().reshape(-1, 2), dim=0)
    m0, m1 = float(mean_val[0]), float(mean_val[1])
    return (int(torch.argmax(mean_val)), m0 if m0 > m1 else abs(1 - m1))


def real_or_fake(prediction):
    return {0: "REAL", 1: "FAKE"}[prediction ^ 1]


def extract_frames(video_file, frames_nums=15):
    from decord import VideoReader, cpu      # host-only dependency, not needed for inference on tensors
    vr = VideoReader(video_file, ctx=cpu(0))
    step = max(1, len(vr) // frames_nums)
    return vr.get_batch(list(range(0, len(vr), step))[:frames_nums]).asnumpy()


def face_rec(frames, p=None, klass=None):
    import cv2
    import dlib
    import face_recognition
    faces = np.zeros((len(frames), 224, 224, 3), dtype=np.uint8)
    count = 0
    mod = "cnn" if dlib.DLIB_USE_CUDA else "hog"
    for frame in frames:
        bgr = cv2.cvtColor(frame, cv2.COLOR_RGB2BGR)
        for top, right, bottom, left in face_recognition.face_locations(bgr, number_of_times_to_upsample=0, model=mod):
            if count >= len(frames):
                break
            crop = cv2.resize(bgr[top:bottom, left:right], (224, 224), interpolation=cv2.INTER_AREA)
            faces[count] = cv2.cvtColor(crop, cv2.COLOR_BGR2RGB)
            count += 1
    return ([], 0) if count == 0 else (faces[:count], count)


def df_face(vid, num_frames, net):
    img = extract_frames(vid, num_frames)
    face, count = face_rec(img)
    return preprocess_frame(face) if count > 0 else []


def is_video(vid):
    return os.path.isfile(vid) and vid.endswith((".avi", ".mp4", ".mpg", ".mpeg", ".mov"))


def set_result():
    return {"video": {"name": [], "pred": [], "klass": [], "pred_label": [], "correct_label": []}}


def store_result(result, filename, y, y_val, klass, correct_label=None, compression=None):
    v = result["video"]
    v["name"].append(filename)
    v["pred"].append(y_val)
    v["klass"].append(klass.lower())
    v["pred_label"].append(real_or_fake(y))
    if correct_label is not None:
        v["correct_label"].append(correct_label)
    if compression is not None:
        v["compression"].append(compression)
    return result
