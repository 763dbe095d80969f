"""CPU oracle for the GenConViT frame-inference forward.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the timed CPU baseline), never as the thing measured or
shipped.  The product path (``model/`` + ``genconvit_b200/``) never imports it
and fails loudly when the CUDA library is missing.

What it is: a functional, fp32, torch-CPU restatement of the reference's
algorithm for the hot path, operating directly on a
``state_dict`` with the reference's key layout:

* ``oracle.nets``       -- GenConViTED / GenConViTVAE / GenConViT forward,
                            pred_vid scoring (reference model/genconvit*.py,
                            model/pred_func.py:111-131)
* ``oracle.backbones``  -- timm==0.6.5 ``convnext_tiny`` and
                            ``swin_tiny_patch4_window7_224`` arithmetic (the
                            un-vendored third-party dependency,
                            requirements.txt:5), restated from the published
                            architecture
* ``oracle.weights``    -- the reference state_dict key/shape inventory and the
                            seeded "trained-like" weight randomiser
* ``oracle.timm_standin`` -- nn.Module shells with timm's module names so the
                            UNMODIFIED reference files import and run here
* ``oracle.make_golden``  -- runs the real reference (from /root/reference, in
                            the authoring container only) and writes
                            ``tests/golden/*.pt``

Pinning (see DESIGN.md "Oracle"): the reference has no tests and no golden
vectors, so the pins are
  (1) backbones vs the independent torchvision implementations
      (``convnext_tiny`` / ``swin_t``), bit-exact under an explicit key map
      (tests/test_oracle.py),
  (2) the restatement vs the reference's own model/*.py executed in the
      authoring container (oracle/make_golden.py -> tests/golden/), and
  (3) the committed golden logits re-checked on every CPU test run.
"""
