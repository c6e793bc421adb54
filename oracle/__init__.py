"""CPU oracle for the U-Net(ResNet-34) hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``vickers_hardness_unet_b200`` (the product) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU baseline.
"""
from .unet_oracle import OracleUnet, OracleDiceLoss, build_oracle, oracle_train_step  # noqa: F401
