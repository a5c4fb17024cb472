"""CPU: the parts of bench.py's contract that do not need a GPU — the reference arm's JSON line (the driver parses it and
computes the speed-up itself) and the pipeline-roofline arithmetic."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--batch", "2", "--phonemes", "8", "--ref-frames", "48", "--cpu-sample-batch", "1"],
                         capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    # "reference" = the unmodified zerovox.tts modules from oracle/_ref (present wherever oracle/build_ref.py ran), else the port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "zerovox", "tts", "model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]") and d["data"] == "synthetic" and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_pipeline_roofline_arithmetic():
    sys.path.insert(0, ROOT)
    import bench
    peaks = {"hbm_gbs": 6500.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1400.0}
    prof = {"gemm_tf32_tcgen05": {"ms": 10.0, "launches": 100, "flops": 3.5e12, "bytes": 1.0e9},      # tensor-bound: 5 ms
            "gemm_3xtf32_tcgen05": {"ms": 2.0, "launches": 30, "flops": 2.3333e11, "bytes": 1.0e8},  # 1 ms at a third of the rate
            "vocoder_conv1d": {"ms": 0.2, "launches": 1, "flops": 1.0e9, "bytes": 6.5e8},            # HBM-bound: 0.1 ms
            "vocoder_upsample": {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0}}
    r = bench.pipeline_roofline(prof, peaks, step_ms=14.0, sm_mhz=1965.0, tf32_tflops=None)
    assert abs(r["t_roof_ms"] - 6.1) < 0.01 and abs(r["t_measured_ms_modelled_classes"] - 12.2) < 1e-9
    assert abs(r["frac"] - 6.1 / 12.2) < 1e-3 and abs(r["share_of_step_modelled"] - 12.2 / 14.0) < 1e-9
    assert r["tf32_peak_tflops"] == 700.0
