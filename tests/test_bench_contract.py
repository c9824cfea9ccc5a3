"""bench.py's contract on the CPU: the reference arm prints exactly one JSON line with the keys the
driver reads, rank != 0 exits without work, and the helpers behind the GPU arm's derived numbers
(algorithmic bytes, published ratio, synthetic particle loads) are right."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True,
                          timeout=600, env=e, cwd=REPO)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run(["--impl", "reference", "--n-parts", "16", "--n-cells", "32", "--steps", "2", "--warmup", "1",
                "--reference-budget-s", "5"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "particle-steps/s" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                      # only 256^3/512^3 has a published number
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    out = _run(["--impl", "reference", "--gpus", "2", "--n-parts", "16", "--n-cells", "32"],
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_and_published_ratio():
    npart, nc = 256 ** 3, 512
    assert bench.b_step_bytes(npart, nc) == 60 * npart + 64 * nc ** 3 == 9596567552      # 9.597 GB
    alg = bench.stage_alg_bytes(npart, nc)
    assert alg["deposit"] == 12 * npart + 4 * nc ** 3 and alg["gather_kick_drift"] == 48 * npart + 4 * nc ** 3
    assert bench.published_ratio(4.7e6, 256, 512) == 1.0 and bench.published_ratio(1.0, 512, 1024) is None


def test_synthetic_particle_loads():
    pos, vel = bench.make_particles_torch(16, 32, "cpu")
    assert tuple(pos.shape) == (3, 4096) and pos.dtype.is_floating_point and float(pos.min()) >= 0 and float(pos.max()) <= 32
    p2, _ = bench.make_particles_torch(16, 32, "cpu")
    assert bool((pos == p2).all())                       # seeded
    # lattice + uniform(-2, 2): every particle within 2 cells of its lattice site (periodic)
    site = (np.arange(16) * 2 + 0.5)
    d = (pos[2].numpy().reshape(16, 16, 16) - site[None, None, :] + 16) % 32 - 16
    assert np.abs(d).max() <= 2.0 + 1e-4
    pc, vc = bench.make_particles_clustered(16, 32)
    cells = np.floor(pc.numpy()).astype(np.int64) % 32
    key = (cells[2] * 32 + cells[1]) * 32 + cells[0]
    assert np.bincount(key).max() > 40                    # the blob: far above the mean of 0.125 per cell
    assert abs(float(vc.std()) - bench.VEL_SIGMA) < 0.01
