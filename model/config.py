"""``load_config()`` -- reference model/config.py:6-10.

The reference opens the cwd-relative ``model/config.yaml`` (and does so at import time
of model/genconvit_vae.py:8).  Here the file next to this module is preferred and the
cwd-relative path is the fallback, so importing the package works from any directory.
"""
import os

import yaml


def load_config():
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.yaml")
    path = here if os.path.exists(here) else os.path.join("model", "config.yaml")
    with open(path) as fh:
        return yaml.safe_load(fh)
