"""bench.py contract checks that need no GPU: the --impl reference arm (CPU oracle) prints exactly one JSON line with
the keys the driver reads, and non-zero ranks of a multi-rank reference launch exit silently."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                           "--warmup", "0", "--cpu-tiles", "4"], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    for k in ["metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"]:
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract_keys():
    """The round's final bench lines under profiles/ (printed by bench.py on the B200 boxes) have every key the driver and the
    judge read, consistent with each other: same metric / workload at every N, value = frames / time, frac = achieved / peak,
    loss and gradient checksums equal across N (the sharded job computes the same step)."""
    lines = {}
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "profiles", f"r2_final_bench_n{n}.json")
        d = json.loads(open(p).read().strip().splitlines()[-1])
        lines[n] = d
        for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "extra"]:
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["unit"] == "frames/s" and d["scaling"] == "strong" and d["data"] == "synthetic"
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
        assert abs(d["value"] - 8 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        r = d["roofline"]
        for k in ["bound", "achieved", "peak", "unit", "frac", "traffic", "issue_frac", "smem_pipe_frac", "frames_per_launch"]:
            assert k in r, (n, k)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1 and 0 < r["issue_frac"] < 1
        assert r["frames_per_launch"] == d["config"]["micro_batch_frames"] == 8 // n
        e = d["e2e"]
        assert e["unit"] == "frames/s" and 0 < e["value"] <= d["value"] * 1.001 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert lines[1]["cpu_baseline"]["kind"] == "port" and lines[1]["cpu_baseline"]["cores"] >= 1
    assert len({d["metric"] for d in lines.values()}) == 1
    ref = lines[1]
    for n, d in lines.items():
        assert abs(d["e2e"]["loss_all_frames"] - ref["e2e"]["loss_all_frames"]) <= 1e-8 * abs(ref["e2e"]["loss_all_frames"])
        for k in ("sum", "l2"):
            assert abs(d["extra"]["grad_checksum"][k] - ref["extra"]["grad_checksum"][k]) <= 1e-6 * abs(ref["extra"]["grad_checksum"][k])
        assert d["value"] > 0.85 * n * ref["value"]  # strong-scaling efficiency of the committed lines
