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


def test_reference_arm_uses_all_host_cores_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the reference on ALL host cores of the
    box whatever launched it (round-1 verdict: the per-N ratios of the scaling record were 8x inflated)."""
    out = _run(["--impl", "reference", "--gpus", "2", "--n-parts", "16", "--n-cells", "32", "--steps", "1", "--warmup", "0",
                "--reference-budget-s", "5"],
               env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == bench.host_threads() >= 1
    if bench.host_threads() > 1:
        assert d["cpu_baseline"]["cores"] > 1


def test_stage_bytes_are_what_each_stage_moves_and_labels_follow_the_sizes():
    """The hand-written solve is five passes over an 8*M-byte array: rows + y forward (stage fft_r2c, 16*M), the
    fused z pass (stage green, 8*M), y + rows inverse (stage fft_c2r, 16*M) -- no stage fraction above 1."""
    npart, nc = 256 ** 3, 512
    m = nc ** 3
    alg = bench.stage_alg_bytes(npart, nc)
    assert alg["green"] == 8 * m and alg["fft_r2c"] == 16 * m and alg["fft_c2r"] == 16 * m
    assert alg["fft_r2c"] + alg["green"] + alg["fft_c2r"] == 40 * m < 56 * m          # the contract charges 56*M
    assert "configs[1]" in bench.workload_label(256, 512) and "configs[3]" in bench.workload_label(1024, 2048)
    assert "configs" not in bench.workload_label(96, 256) and "96^3 particles on 256^3 mesh" in bench.workload_label(96, 256)
    assert "configs[4]" in bench.workload_label(256, 512, "evolved") and "configs[4]" in bench.workload_label(256, 512, "clustered")


def test_weak_scaling_sub_record_is_wired_to_the_eight_gpu_line():
    """--gpus 8 on the default workload adds `weak_scaling` (configs[3] on 8 GPUs against configs[2] on one) to
    the one JSON line; the record's keys are what the judge reads.  Checked on the source (no GPU here); the
    8-GPU run of round 2 is profiles/r02_r8_bench_g8_default.json."""
    import inspect
    src = inspect.getsource(bench.run_slab)
    assert "world == 8 and (n_parts, n_cells) == (256, 512)" in src and '"weak_scaling": weak' in src
    rec_src = inspect.getsource(bench.weak_scaling_record)
    for k in ["ms_per_step_1gpu", "ms_per_step_all_gpus", "parallel_efficiency", "target", "phases_ms_rank0",
              "workload_1gpu", "workload_all_gpus", "roofline_step"]:
        assert k in rec_src, k
    path = os.path.join(REPO, "profiles", "r02_r8_bench_g8_default.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    w = d["weak_scaling"]
    assert d["n_gpus"] == 8 and w["n_gpus"] == 8 and "configs[2]" in w["workload_1gpu"] and "configs[3]" in w["workload_all_gpus"]
    assert abs(w["parallel_efficiency"] - w["ms_per_step_1gpu"] / w["ms_per_step_all_gpus"]) < 1e-12
