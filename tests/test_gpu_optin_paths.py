"""The opt-in kernel paths stay parity-tested by the regular GPU run: both knobs are read once per process, so each mode
re-runs the relevant parity tests in a child process with the knob set.
  PLK_NTT_TMA=1              every qualifying NTT pass through the TMA / Stockham kernel (csrc/ntt_tma.cuh)
  PLK_MSM_AFFINE_ROUNDS=2    batched-affine bucket rounds in front of the XYZZ accumulation (csrc/msm_affine.cuh)
  PLK_MSM_MADD_COMPACT=0/1/2 the other code-size variants of the mixed addition (3 = four products as calls is the default)
  PLK_MSM_BATCH_MERGE=16     batches of short vectors as ONE merged pipeline with a bucket set per vector (default: fork/join)
  PLK_MSM_SCATTER_PASS_KB=16 the range-partitioned scatter (default only beyond 96 MiB of sorted entries) forced onto small inputs
  PLK_MSM_OVERLAP_PARTS=2/8  the overlapped pipeline: bucket ranges accumulated on their own streams, reduction tails underneath"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_child(env_extra, select, files):
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", select, "-p", "no:cacheprovider"] + [os.path.join(ROOT, "tests", f) for f in files]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_ntt_parity_with_every_pass_through_tma():
    run_child({"PLK_NTT_TMA": "1"}, "fft or ntt or lde or coset or divide or poly or domain", ["test_gpu_parity.py", "test_gpu_edge.py"])


def test_ntt_parity_with_on_the_fly_twiddles_only():
    run_child({"PLK_NTT_NO_DIRECT": "1"}, "fft or ntt or lde or coset or divide", ["test_gpu_parity.py"])


@pytest.mark.parametrize("rounds", ["1", "2", "3"])
def test_msm_parity_with_batched_affine_rounds(rounds):
    run_child({"PLK_MSM_AFFINE_ROUNDS": rounds}, "msm or ipa or summ", ["test_gpu_parity.py", "test_gpu_edge.py", "test_gpu_sharded.py"])


@pytest.mark.parametrize("variant", ["0", "1", "2"])
def test_msm_parity_with_other_madd_variants(variant):
    run_child({"PLK_MSM_MADD_COMPACT": variant}, "msm", ["test_gpu_parity.py", "test_gpu_edge.py"])


@pytest.mark.parametrize("parts,mode", [("2", "0"), ("8", "0"), ("4", "1")])
def test_msm_parity_with_overlapped_pipeline(parts, mode):
    run_child({"PLK_MSM_OVERLAP_PARTS": parts, "PLK_MSM_OVERLAP_MODE": mode}, "msm or shard or ipa", ["test_gpu_parity.py", "test_gpu_edge.py", "test_gpu_sharded.py"])


@pytest.mark.parametrize("width", ["16", "3"])
def test_msm_batches_as_merged_pipeline(width):
    run_child({"PLK_MSM_BATCH_MERGE": width}, "batch or commit or ipa", ["test_gpu_parity.py", "test_serde.py", "test_ipa.py"])


def test_msm_parity_with_range_partitioned_scatter():
    run_child({"PLK_MSM_SCATTER_PASS_KB": "16"}, "msm or shard", ["test_gpu_parity.py", "test_gpu_edge.py", "test_gpu_sharded.py"])
