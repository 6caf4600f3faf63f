"""bench.py's CPU-side contract: the reference arm runs without a GPU and prints one JSON line with the keys the
driver reads; under a multi-rank launch only rank 0 prints; the default arm refuses to run without a CUDA device
(no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("SNN-head images/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["T_rpn"] == 8 and d["config"]["T_det"] == 12


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_canonical_flop_count_is_the_surveys():
    """the reference's own FLOP count per image (every layer at all T steps, 1 MAC = 2 FLOP):
    Cityscapes T 8/12, 9 classes = 1267.4 GFLOP; BDD, 5 classes = 1169.8 GFLOP."""
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.canonical_gflop_per_image("cityscapes", 8, 12) - 1267.4) < 0.05
    assert abs(bench.canonical_gflop_per_image("bdd", 8, 12) - 1169.8) < 0.05
    a = bench.parse([])
    assert a.steps == 100 and a.warmup == 10 and a.gpus == 1 and a.mode == "fp16x2" and a.precondition_s == 1.0


def test_bench_inputs_are_seeded_per_global_image_index():
    """shards are reproducible because every image is drawn from its own seed, whichever rank owns it."""
    sys.path.insert(0, ROOT)
    import torch
    import bench
    f0, r0 = bench.bench_inputs("bdd", 3)
    f1, r1 = bench.bench_inputs("bdd", 3)
    f2, _ = bench.bench_inputs("bdd", 4)
    assert all(torch.equal(a, b) for a, b in zip(f0, f1)) and torch.equal(r0, r1)
    assert not torch.equal(f0[4], f2[4])
    assert [tuple(f.shape) for f in f0] == [(256, h, w) for (h, w) in bench.BDD_LEVELS] and r0.shape == (1000, 256, 7, 7)
