"""Host-side logic that needs no GPU: synthetic workload generator, image sharding, and the world_size-2 gloo
gather used by the multi-GPU bench (no collective on the data path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from multiposenet_b200 import parallel, synthetic  # noqa: E402


def test_workload_shapes_match_baseline():
    assert synthetic.WORKLOADS["c1"].num_anchors == 32736
    assert synthetic.WORKLOADS["c2"].num_anchors == 76725 and synthetic.WORKLOADS["c2"].n_loc == 9
    assert synthetic.WORKLOADS["c2_n6"].num_anchors == 51150
    assert synthetic.WORKLOADS["c4"].num_anchors == 130944
    assert synthetic.seed_for(2, 3) == 20240203


def test_generator_is_deterministic_and_guard_banded():
    wl = synthetic.WORKLOADS["tiny"]
    a = synthetic.make_inputs(wl)
    b = synthetic.make_inputs(wl)
    for k in ("class_logits", "encoded_boxes", "heatmap_logits"):
        assert a[k].dtype == np.float32 and np.array_equal(a[k], b[k])
    A = wl.num_anchors
    assert a["class_logits"].shape == (wl.batch, A) and a["encoded_boxes"].shape == (wl.batch, A, 4)
    assert a["heatmap_logits"].shape == (wl.batch, wl.height // 4, wl.width // 4, 18)
    assert np.abs(a["encoded_boxes"]).max() <= 4.0
    lt = np.log(0.3 / 0.7)
    assert np.abs(a["class_logits"] - lt).min() > 5e-4                  # nothing near the score threshold
    for row in a["class_logits"]:
        conf = np.sort(row[row > lt])
        assert conf.size >= 4 and np.diff(conf).min() > 5e-5            # confident logits are tie-free
    # every keypoint channel has a real peak, so M != m (no 0/0 in the normalisation)
    hm = a["heatmap_logits"][..., :17]
    assert (hm.max(axis=(1, 2)) > 2.0).all()
    c = synthetic.make_inputs(wl, replicate=1)
    assert not np.array_equal(a["class_logits"], c["class_logits"])


def test_anchor_placement_agrees_with_oracle():
    import oracle
    for key in ("tiny", "c2"):
        wl = synthetic.WORKLOADS[key]
        a = synthetic.anchors_np(wl.height, wl.width, wl.strides, wl.scales, wl.multipliers, wl.ratios)
        b = oracle.anchors(wl.height, wl.width, wl.strides, wl.scales, wl.multipliers, wl.ratios)
        np.testing.assert_allclose(a, b, atol=1e-6)


def test_prn_weight_init_statistics():
    W1, b1, W2, b2 = synthetic.make_prn_weights(d=17 * 8 * 8, hidden=64)
    assert W1.shape == (1088, 64) and W2.shape == (64, 1088) and not b1.any() and not b2.any()
    s1 = np.sqrt(1.0 / 1088) / 0.87962566
    assert np.abs(W1).max() <= 2 * s1 * 1.0001 and abs(W1.std() / (s1 * 0.87962566) - 1) < 0.05


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _fake_result(lo, hi, max_det=5):
    rng = np.random.default_rng(100 + lo)
    B = hi - lo
    num = rng.integers(0, max_det + 1, B).astype(np.int32)
    N = int(num.sum())
    return {"boxes": rng.random((B, max_det, 4)).astype(np.float32), "scores": rng.random((B, max_det)).astype(np.float32),
            "num_boxes": num, "keypoint_scores": np.full((N, 17), float(lo), np.float32),
            "keypoint_positions": np.zeros((N, 17, 2), np.float32)}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(n_images, rank, world)
        res = _fake_result(lo, hi)
        merged = parallel.gather_results(res, dst=0)
        # the fixed-size gather of the bench's e2e leg: padded per-call buffers (persons padded to B * max_det rows)
        lo0, hi0 = parallel.shard_range(8, rank, world)               # equal shards: every rank's block has the same size
        pad = _fake_result(lo0, hi0)
        n = int(pad["num_boxes"].sum())
        rows = (hi0 - lo0) * 5
        pad["keypoint_scores"] = np.concatenate([pad["keypoint_scores"], np.full((rows - n, 17), -1, np.float32)])
        pad["keypoint_positions"] = np.concatenate([pad["keypoint_positions"], np.full((rows - n, 17, 2), -1, np.float32)])
        packed = parallel.gather_packed({k: torch.from_numpy(v) for k, v in pad.items()}, dst=0)
        if rank == 0:
            want = parallel.merge_results([_fake_result(*parallel.shard_range(8, r, world)) for r in range(world)])
            for k in want:
                assert np.array_equal(packed[k], want[k]), k
        else:
            assert packed is None
        # the bench's timing reduction: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        if rank == 0:
            np.savez(out_path, **merged)
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo_gather(tmp_path):
    n_images, world = 7, 2
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), n_images, out), nprocs=world, join=True)
    m = np.load(out)
    parts = [_fake_result(*parallel.shard_range(n_images, r, world)) for r in range(world)]
    assert np.array_equal(m["num_boxes"], np.concatenate([p["num_boxes"] for p in parts]))
    assert np.array_equal(m["boxes"], np.concatenate([p["boxes"] for p in parts]))
    assert m["person_offsets"][-1] == m["keypoint_scores"].shape[0] == sum(int(p["num_boxes"].sum()) for p in parts)
    assert np.array_equal(np.diff(m["person_offsets"]), m["num_boxes"])
    # persons stay in image order: the marker written by rank 1 comes after rank 0's rows
    n0 = int(parts[0]["num_boxes"].sum())
    assert (m["keypoint_scores"][:n0] == 0.0).all() and (m["keypoint_scores"][n0:] == 4.0).all()


def test_bench_and_tools_compile_and_bench_formulas_hold():
    """bench.py and the development tools at least parse; the algorithmic byte counts of DESIGN.md (section 4) are what
    bench.py's roofline uses."""
    import glob
    import py_compile
    for path in [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
            sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))):
        py_compile.compile(path, doraise=True)
    sys.path.insert(0, ROOT)
    import bench
    wl = synthetic.WORKLOADS["c2"]
    D, Hd = 56 * 36 * 17, 1024
    # (compulsory per SURVEY 8(d), moved by the kernel as built)
    w = 2 * D * Hd * 2 + (Hd + D) * 4
    assert w == 140378112 + 141184
    assert bench.algorithmic_bytes("prn_fused", wl, 8, 77, 1825, "bf16") == (w + 77 * 274176, w + 77 * D * 10)
    assert bench.algorithmic_bytes("prn_fused", wl, 32, 2805, 66877, "bf16") == (None, None)      # idle above 256 persons
    assert bench.algorithmic_bytes("heatmap", wl, 8, 77, 1825, "bf16")[0] == 2 * 72 * 160 * 160 * 8
    assert bench.algorithmic_bytes("heatmap_norm", wl, 8, 77, 1825, "bf16") == (144 * 25600 * 8, 224 * 25600 * 8)
    assert bench.algorithmic_bytes("logit_minmax", wl, 8, 77, 1825, "bf16") == (None, 72 * 25600 * 8)
    assert bench.algorithmic_bytes("keypoint_decode", wl, 8, 77, 1825, "bf16")[0] == 77 * D * 4
    # both arms describe the workload identically (the driver compares `config`), labelled from the workload itself
    import argparse
    a = argparse.Namespace(prn_mode="bf16", gpus=1, scaling="weak", lanes=3)
    assert bench.config_dict(wl, a) == bench.config_dict(wl, argparse.Namespace(prn_mode="bf16", gpus=1, scaling="weak"))
    assert bench.config_dict(synthetic.WORKLOADS["c4"], a)["workload"].startswith("BASELINE configs[3]")
    assert bench.batch_per_gpu(synthetic.WORKLOADS["c4"], argparse.Namespace(scaling="strong", gpus=8)) == 8


@pytest.mark.timeout(600)
def test_reference_arm_prints_exactly_one_json_line_on_stdout():
    """`bench.py --impl reference` (the CPU arm: oracle port on the host cores, no GPU involved): stdout carries ONE line,
    it parses, and it has the keys of the bench contract; whatever else the process prints goes to stderr."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("BASELINE configs[1]")


def test_lanes_argument_is_checked_before_anything_touches_a_device():
    from multiposenet_b200 import DetectorLanes
    with pytest.raises(ValueError):
        DetectorLanes(None, lanes=0)


def test_importing_bench_leaves_stdout_alone():
    sys.path.insert(0, ROOT)
    import bench
    assert bench._REAL_STDOUT is None
