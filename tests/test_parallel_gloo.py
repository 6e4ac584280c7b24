"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray/frame sharding, the final image gather and the
gradient all-reduce.  The kernels are rank-local; tests/test_gpu_parity.py::test_full_size_properties_c2 checks on
the GPU that a slice rendered alone equals the same slice of the full frame."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quadraturefields_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        W, H = 24, 20
        N = W * H
        frame = torch.arange(N * 5, dtype=torch.float32).reshape(N, 5)          # stands for (rgb, alpha, depth) of the frame
        lo, hi = P.shard_rays(N, rank, ws, image_width=W)
        sizes = [b - a for a, b in (P.shard_rays(N, r, ws, image_width=W) for r in range(ws))]
        got = P.gather_frame(frame[lo:hi] * 1.0, sizes, dst=0)
        ok = True
        if rank == 0:
            ok &= torch.equal(got, frame)
        else:
            ok &= got is None
        # gradient all-reduce: per-rank mean gradients, unequal sample counts -> global mean
        p = torch.nn.Parameter(torch.zeros(7))
        n_local = 10 + 30 * rank
        p.grad = torch.full((7,), float(rank + 1))
        P.all_reduce_gradients([p], n_local)
        expect = sum((10 + 30 * r) * (r + 1) for r in range(ws)) / sum(10 + 30 * r for r in range(ws))
        ok &= bool(torch.allclose(p.grad, torch.full((7,), expect)))
        # a rank whose batch hit nothing has no .grad at all: it must still enter the collectives (zeros), not return early
        big = torch.nn.Parameter(torch.zeros(1 << 20))
        small, unused = torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(0))
        if rank == 1:
            big.grad, small.grad = torch.full_like(big, 3.0), torch.full_like(small, 5.0)
        n_loc = torch.tensor(0.0 if rank == 0 else 8.0)                 # device-side count: no host read of the total
        P.all_reduce_gradients([big, small, unused], n_loc)
        ok &= bool(torch.allclose(big.grad, torch.full_like(big, 3.0))) and bool(torch.allclose(small.grad, torch.full_like(small, 5.0)))
        ok &= unused.grad is None
        ok &= P.max_over_ranks(1.0 + rank) == float(ws)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ws = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(ws, _free_port(), ret), nprocs=ws, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_sharding_covers_every_ray_once():
    for n, w in ((640000, 800), (2073600, 1920), (1000, 0), (10, 0), (7, 0)):
        for ws in (1, 2, 3, 4, 8):
            cuts = [P.shard_rays(n, r, ws, w) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(ws - 1))
            if w:
                assert all((b - a) % (4 * w) == 0 for a, b in cuts)
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= (4 * w if w else 1)
    views = sorted(P.view_for_step(s, r, 4, 200) for s in range(50) for r in range(4))
    assert views == list(range(200))


def test_band_cyclic_split_and_assembly():
    """Band-cyclic frame sharding (bench.py c4 / c5 legs): every row belongs to exactly one rank, shares reassemble."""
    for H, W in ((1080, 16), (2160, 8), (48, 48), (100, 4)):
        rows = torch.arange(H).view(H, 1).expand(H, W).reshape(-1, 1).float()
        for ws in (1, 2, 3, 4, 8):
            sizes = [P.band_rows(H, r, ws) for r in range(ws)]
            assert sum(sizes) == H and max(sizes) - min(sizes) <= 4
            parts = []
            for r in range(ws):
                mine = [y for y in range(H) if (y // 4) % ws == r]
                assert len(mine) == sizes[r]
                parts.append(rows.view(H, W)[mine].reshape(-1, 1))
            assert torch.equal(P.assemble_banded(parts, H, W), rows)


def test_band_cyclic_split_with_other_band_heights():
    """`render_pose(band_rows=...)` / `qf_band_rows` / `assemble_banded(band=...)` agree for bands thicker than 4 rows (the
    C ABI's row count is the host helper's; the library is loaded, no kernel runs), and the whole-frame row of a share's local
    row — the map `composite_rays_kernel` applies for `qf_render_mesh_*_to_frame` — is a bijection onto the frame."""
    from quadraturefields_b200 import _lib
    lib = _lib.load()
    for H, W, band in ((1080, 8, 12), (1080, 8, 40), (2160, 8, 16), (96, 4, 8)):
        rows = torch.arange(H).view(H, 1).expand(H, W).reshape(-1, 1).float()
        for ws in (1, 2, 3, 5, 8):
            parts, seen = [], []
            for r in range(ws):
                mine = [y for y in range(H) if (y // band) % ws == r]
                assert P.band_rows(H, r, ws, band) == len(mine) == int(lib.qf_band_rows(H, band, ws, r))
                parts.append(rows.view(H, W)[mine].reshape(-1, 1))
                # FrameMap of render.cu: local row -> whole-frame row
                seen += [((l // band) * ws + r) * band + l % band for l in range(len(mine))]
                assert seen[-len(mine):] == mine
            assert sorted(seen) == list(range(H))
            assert torch.equal(P.assemble_banded(parts, H, W, band), rows)
    assert int(lib.qf_band_rows(1080, 4, 8, 8)) < 0          # rank outside the world


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` launched the way the driver launches N>1 arms: rank 0 alone measures and prints the
    JSON line (impl, metric, cpu_baseline, e2e with zero copy bytes), the other rank exits 0 without work."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rays_per_sec_render_fwd" and d["unit"] == "rays/s" and d["n_gpus"] == 2
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "BVH" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
