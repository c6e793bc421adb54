"""Native kernel-level parity (B200): build/selftest compares every conv / pool / head / weight-gradient kernel of
csrc/ against a CPU reference written in the same file (tests/native/selftest.cu), including the edge shapes the Python
tests do not reach (partial tiles, multi-tile-per-CTA, narrow images, odd tile counts in CTA pairs)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "selftest")
BIN_LIB = os.path.join(ROOT, "build", "selftest_lib")   # no -DUB_TC_PROF: the kernel binaries of libunetb200.so

# group filter -> what it covers (a filter selects every group whose name contains it)
GROUPS = {
    "pool": "max-pool 3x3 s2",
    "head": "CUDA-core seg head (reference kernel)",
    "conv": "tap-table igemm, cp.async halo conv (hconv), TMA halo conv incl. parity mode (tconv)",
    "stem": "7x7 s2 stem: tap-table and overlapped-row tconv mode",
    "dec1": "two-source parity decoder conv1 (igemm)",
    "wide": "halo-resident wide conv with streamed weights (wconv)",
    "split": "wide decoder conv1 as wpconv + wconv with residual",
    "pair": "CTA-pair (cta_group::2) wide conv (wconv2)",
    "hwgrad": "halo-resident weight gradient",
    "swgrad": "7x7/s2 stem weight gradient (input-row anchors, two dZ rows per MMA)",
    "dlow": "decoder conv1 data gradient w.r.t. the low-resolution input (parity images + 16 shifted taps)",
    "xwgrad": "TMA-fed halo weight gradient (all nine taps per pass; narrow and wide modes)",
}


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_native_selftest(group):
    if not os.path.exists(BIN):
        pytest.fail(f"{BIN} is missing: run `make` (the GPU box receives the built binary with the repo snapshot)")
    r = subprocess.run([BIN, group], capture_output=True, text=True, timeout=300, cwd=ROOT)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0 and "SELFTEST OK" in r.stdout and "[FAIL]" not in r.stdout, tail
    n = int(r.stdout.rsplit("SELFTEST OK:", 1)[1].split("checks")[0])
    assert n > 0, f"filter {group!r} selected no checks"
    print(f"{group}: {n} checks ({GROUPS[group]})")


@pytest.mark.parametrize("group", ["conv", "stem", "wide", "split", "xwgrad", "dlow", "swgrad"])
def test_native_selftest_library_build(group):
    """The same checks on the binaries the LIBRARY contains: build/selftest is compiled with -DUB_TC_PROF (role-cycle
    counters change register allocation of the tconv kernels), build/selftest_lib is not."""
    if not os.path.exists(BIN_LIB):
        pytest.fail(f"{BIN_LIB} is missing: run `make`")
    r = subprocess.run([BIN_LIB, group], capture_output=True, text=True, timeout=300, cwd=ROOT)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0 and "SELFTEST OK" in r.stdout and "[FAIL]" not in r.stdout, tail
    assert int(r.stdout.rsplit("SELFTEST OK:", 1)[1].split("checks")[0]) > 0
