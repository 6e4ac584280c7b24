"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

It puts ``oracle/shims`` (stand-ins for the absent third-party imports) and
``/root/reference/examples`` on ``sys.path``, neutralises ``.cuda()`` / ``device="cuda"`` so the
CUDA-only reference code executes on CPU, imports the reference modules as they are, calls the
hot-path functions on seeded inputs and stores inputs + outputs.  The fixtures pin
``oracle/quadfield_oracle.py`` (tests/test_oracle_golden.py) and, through it, the CUDA kernels.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("QF_REFERENCE", "/root/reference/examples")
OUT = os.path.join(ROOT, "tests", "golden")


def _install():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    warnings.filterwarnings("ignore")
    # CUDA-only reference code → CPU
    torch.Tensor.cuda = lambda self, *a, **k: self

    def _strip(fn):
        def wrapped(*a, **k):
            dev = k.get("device", None)
            if dev is not None and "cuda" in str(dev):
                k.pop("device")
            return fn(*a, **k)
        return wrapped

    for name in ("zeros", "ones", "zeros_like", "ones_like", "rand", "empty", "tensor", "full"):
        setattr(torch, name, _strip(getattr(torch, name)))


def golden_field_rendering():
    """a12–a15: reference field_rendering.py (behind the nerfacc pack/scan stand-in)."""
    import field_rendering as FR
    g = torch.Generator().manual_seed(42)
    counts = torch.tensor([3, 0, 5, 1, 0, 0, 8, 2, 4, 1, 7, 0], dtype=torch.long)
    n_rays = counts.numel()
    ray_indices = torch.repeat_interleave(torch.arange(n_rays), counts)
    M = int(counts.sum())
    alphas = torch.rand(M, generator=g)
    alphas[5] = 0.0
    alphas[9] = 1.0
    sigmas = torch.rand(M, generator=g) * 40
    t_starts = torch.rand(M, generator=g).cumsum(0) * 0.05
    t_ends = t_starts + 0.02 + 0.01 * torch.rand(M, generator=g)
    rgbs = torch.rand(M, 3, generator=g)
    prefix = torch.rand(M, generator=g)
    bkgd = torch.tensor([0.2, 0.5, 0.9])
    out = dict(counts=counts, ray_indices=ray_indices, alphas=alphas, sigmas=sigmas, t_starts=t_starts,
               t_ends=t_ends, rgbs=rgbs, prefix=prefix, bkgd=bkgd)
    out["T_alpha"] = FR.render_transmittance_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays)
    out["T_alpha_prefix"] = FR.render_transmittance_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays,
                                                               prefix_trans=prefix)
    w, T = FR.render_weight_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays)
    out["w_alpha"], out["T_alpha2"] = w, T
    w, T, a = FR.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=n_rays)
    out["w_density"], out["T_density"], out["a_density"] = w, T, a
    w, T, a = FR.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=n_rays,
                                            prefix_trans=prefix)
    out["w_density_prefix"] = w
    out["vis_alpha"] = FR.render_visibility_from_alpha(alphas, ray_indices=ray_indices, n_rays=n_rays,
                                                       early_stop_eps=0.3, alpha_thre=0.2)
    out["vis_density"] = FR.render_visibility_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices,
                                                           n_rays=n_rays, early_stop_eps=0.05, alpha_thre=0.3)
    out["acc_rgb"] = FR.accumulate_along_rays(out["w_density"], rgbs, ray_indices, n_rays)
    out["acc_w"] = FR.accumulate_along_rays(out["w_density"], None, ray_indices, n_rays)
    c, o, d, ex = FR.rendering(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn=lambda a, b, c: (rgbs, sigmas),
                               render_bkgd=bkgd)
    out["rend_c"], out["rend_o"], out["rend_d"] = c, o, d
    c, o, d, ex = FR.rendering(t_starts, t_ends, ray_indices, n_rays, rgb_alpha_fn=lambda a, b, c: (rgbs, alphas))
    out["renda_c"], out["renda_o"], out["renda_d"] = c, o, d
    c, o, d, w, wr = FR.rendering_field(t_starts, t_ends, ray_indices, n_rays,
                                        rgb_sigma_fn=lambda a, b, c: (rgbs, sigmas), render_bkgd=bkgd)
    out["rf_c"], out["rf_o"], out["rf_d"], out["rf_w"], out["rf_wrev"] = c, o, d, w, wr
    # batched layout (n_rays, S)
    ab = torch.rand(6, 9, generator=g)
    sb = torch.rand(6, 9, generator=g) * 30
    tsb = torch.rand(6, 9, generator=g).cumsum(1) * 0.05
    teb = tsb + 0.02
    vb = torch.rand(6, 9, 3, generator=g)
    out.update(b_alphas=ab, b_sigmas=sb, b_ts=tsb, b_te=teb, b_vals=vb)
    out["b_w_alpha"], out["b_T_alpha"] = FR.render_weight_from_alpha(ab)
    out["b_w_density"], out["b_T_density"], out["b_a_density"] = FR.render_weight_from_density(tsb, teb, sb)
    out["b_acc"] = FR.accumulate_along_rays(out["b_w_density"], vb)
    # docstring KATs (field_rendering.py:192-195, 246-253, 298-302, 347-355, 403-409)
    ka = torch.tensor([0.4, 0.8, 0.1, 0.8, 0.1, 0.0, 0.9])
    kr = torch.tensor([0, 0, 0, 1, 1, 2, 2])
    out["kat_T"] = FR.render_transmittance_from_alpha(ka, ray_indices=kr)
    out["kat_w"], _ = FR.render_weight_from_alpha(ka, ray_indices=kr)
    kts, kte = torch.arange(7.0), torch.arange(7.0) + 1
    out["kat_wd"], out["kat_Td"], out["kat_ad"] = FR.render_weight_from_density(kts, kte, ka, ray_indices=kr)
    out["kat_vis"] = FR.render_visibility_from_alpha(ka, ray_indices=kr, early_stop_eps=0.3, alpha_thre=0.2)
    return out


def golden_sg_decode():
    """a7, a8, a9: reference texture_utils.FeatureCompression / ngp.py dequantisers / SG head / _TruncExp."""
    import radiance_fields.ngp as NGP
    from texture_utils import FeatureCompression
    from oracle import quadfield_oracle as O
    out = {}
    for L, ctype, lam in ((3, "linear", 5.0), (6, "sigmoid", 7.5), (2, "sigma", 7.5)):
        S = 24
        tex = O.make_texture_set(S, L, seed=100 + L, compression_type=ctype, lambda_thres=lam)
        fc = object.__new__(FeatureCompression)
        fc.num_lobes, fc.texture_size, fc.compression_type, fc.lambda_thres = L, S, ctype, lam
        fc.alpha, fc.diffuse = tex.alpha, tex.diffuse
        fc.sg_colors = {i: tex.sg_colors[i] for i in range(L)}
        fc.lambdas = {i: tex.lambdas[i] for i in range(L)}
        g = torch.Generator().manual_seed(7 + L)
        idx = torch.randint(0, S, (300, 2), generator=g)
        # make sure extreme quantised values are exercised
        tex.alpha[idx[0, 0], idx[0, 1]] = 255
        tex.alpha[idx[1, 0], idx[1, 1]] = 0
        feats = fc.get_features_from_texture_map(idx)
        dirs = torch.nn.functional.normalize(torch.randn(300, 3, generator=g), dim=-1)
        sg = object.__new__(NGP.NGPRadianceFieldSGNew)
        torch.nn.Module.__init__(sg)
        sg.num_g_lobes, sg.discretize = L, False
        rgb = sg.features_to_rgb(feats[:, :-1], dirs)
        rgb2 = fc.features_to_rgb(feats[:, :-1], dirs)
        assert torch.equal(rgb, rgb2)
        k = f"L{L}_{ctype}"
        out[k + "_idx"], out[k + "_feats"], out[k + "_dirs"], out[k + "_rgb"] = idx, feats, dirs, rgb
        out[k + "_alpha"], out[k + "_diffuse"] = tex.alpha, tex.diffuse
        for i in range(L):
            out[k + f"_color{i}"], out[k + f"_lambda{i}"] = tex.sg_colors[i], tex.lambdas[i]
    # f-4: the bake writer, FeatureCompression.compress (texture_utils.py:67-98) on random features
    for L, ctype, lam in ((3, "linear", 5.0), (2, "sigma", 7.5)):
        g = torch.Generator().manual_seed(50 + L)
        M = 400
        feats = torch.randn(M, 3 + 7 * L + 1, generator=g) * 3.0
        feats[:, -1] = torch.rand(M, generator=g) * 600          # sigma >= 0
        feats[:5, -1] = torch.tensor([0.0, 1e-3, 50.0, 5000.0, 1e6])
        fc = object.__new__(FeatureCompression)
        fc.num_lobes, fc.compression_type, fc.lambda_thres = L, ctype, lam
        data = fc.compress(feats)
        k = f"cmp_L{L}_{ctype}"
        out[k + "_feats"], out[k + "_alpha"], out[k + "_diffuse"] = feats, data["alpha"], data["diffuse"]
        for i in range(L):
            out[k + f"_lambda{i}"], out[k + f"_color{i}"] = data["lambdas"][i], data["colors"][i]
    x = torch.tensor([-3.0, -1.0, 0.0, 0.5, 2.0, 14.0, 16.0, 20.0], requires_grad=True)
    y = NGP.trunc_exp(x)
    y.backward(torch.ones_like(y))
    out["trunc_exp_x"], out["trunc_exp_y"], out["trunc_exp_g"] = x.detach(), y.detach(), x.grad
    return out


def golden_geometry():
    """a2 (plane-hit formula), a3, a4: reference mesh_utils functions; the intersector itself
    (Embree/OptiX, absent) is replaced by the oracle's brute-force `intersects_id`."""
    import mesh_utils as MU
    import trimesh
    from oracle import quadfield_oracle as O
    verts, faces = O.shell_mesh([0.5, 0.8, 1.0], subdivisions=2, jitter=1e-3, seed=3)
    f, cx, cy, W, H = O.pinhole_intrinsics(24, 24, 0.6911)
    c2w = O.look_at_c2w((2.4, 1.9, 1.3))
    origins, viewdirs = O.generate_rays(c2w, W, H, f, cx, cy)
    K = 4
    mesh = trimesh.Trimesh(vertices=verts, faces=faces, process=False)

    class FakeIntersector:
        def intersects_id(self, o, v, multiple_hits=True, return_locations=True, max_hits=10):
            return O.intersects_id(o, v, verts, faces, max_hits)

    mi = object.__new__(MU.MeshIntersection)
    mi.mesh, mi.num_intersections, mi.render_step_size = mesh, K, 0.005
    mi.rayintersector = FakeIntersector()
    tup = mi.sampling_raytrace_numpy(viewdirs, origins)
    points, vectors, index_ray, depth, index_tri, _, org = tup
    out = dict(verts=verts, faces=faces, origins=origins, viewdirs=viewdirs, K=np.int64(K),
               points=points, vectors=vectors, index_ray=index_ray, depth=depth, index_tri=index_tri, org=org)
    # reference plane-hit formula on its own (jit function, mesh_utils.py:33-40)
    n = torch.from_numpy(mesh.face_normals[index_tri].astype(np.float32))
    v = torch.from_numpy(mesh.vertices[mesh.faces[index_tri][:, 0]].astype(np.float32))
    psi = MU.ray_triangle_intersection(torch.from_numpy(origins[index_ray]), torch.from_numpy(viewdirs[index_ray]), n, v)
    out["psi"] = psi
    # a4: perturb depths so the re-sort actually permutes, then reference sampling_indexing
    g = torch.Generator().manual_seed(5)
    d2 = torch.from_numpy(depth.astype(np.float32)) + 0.3 * torch.randn(depth.shape[0], generator=g)
    res = mi.sampling_indexing(torch.from_numpy(points.astype(np.float32)), torch.from_numpy(org.astype(np.float32)),
                               torch.from_numpy(vectors.astype(np.float32)), torch.from_numpy(index_ray),
                               d2, torch.from_numpy(index_tri))
    names = ("points", "deltas", "boundary", "vectors", "index_ray", "depth", "index_tri", "origins")
    out["si_in_depth"] = d2
    for nme, r in zip(names, res):
        out["si_" + nme] = r
    return out


def golden_derive_properties():
    """a11: reference utils.derive_properties (kaolin pack scans behind the stand-in)."""
    import utils as U
    g = torch.Generator().manual_seed(11)
    counts = torch.tensor([2, 0, 4, 1, 0, 8, 3, 0], dtype=torch.long)
    N = counts.numel()
    index_ray = torch.repeat_interleave(torch.arange(N), counts)
    M = int(counts.sum())
    color = torch.rand(M, 3, generator=g)
    density = torch.rand(M, generator=g) * 300
    depths = torch.rand(M, generator=g) * 5
    deltas = torch.full((M,), 0.005)
    boundary = torch.ones(M, dtype=torch.bool)
    boundary[1:] = index_ray[1:] != index_ray[:-1]
    bk = torch.tensor([0.1, 0.6, 0.3])
    out = dict(counts=counts, index_ray=index_ray, color=color, density=density, depths=depths, deltas=deltas,
               boundary=boundary, bk=bk)
    for bg in ("white", "black", "random"):
        rgb, a, ids, D, w = U.derive_properties(color, density, depths, deltas, boundary, index_ray,
                                                render_bkgd=bk, bg_color=bg, N=N)
        out[bg + "_rgb"], out[bg + "_alpha"], out[bg + "_ids"], out[bg + "_depth"], out[bg + "_w"] = rgb, a, ids, D, w
    return out


def golden_ngp():
    """a5, a6: reference NGPRadianceField module glue over the tinycudann stand-in."""
    import radiance_fields.ngp as NGP
    from oracle import quadfield_oracle as O
    log2_T = 12
    p = O.make_ngp_params(seed=42, log2_hashmap_size=log2_T, table_scale=2e3)
    rf = NGP.NGPRadianceField(aabb=p.aabb.tolist(), log2_hashmap_size=log2_T)
    with torch.no_grad():
        rf.mlp_base.params.copy_(torch.cat([w.flatten() for w in p.base_w] + [p.table.flatten()]))
        rf.mlp_head.params.copy_(torch.cat([w.flatten() for w in p.head_w]))
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(400, 3, generator=g) * 2 - 1) * 1.45
    x[:8] *= 1.2  # a few points outside the aabb (selector path)
    d = torch.nn.functional.normalize(torch.randn(400, 3, generator=g), dim=-1)
    rgb, density = rf(x, d)
    dens2, feat = rf.query_density(x, return_feat=True)
    sel, xn = rf.normalize(x)
    return dict(log2_T=np.int64(log2_T), seed=np.int64(42), table_scale=np.float64(2e3), x=x, d=d,
                rgb=rgb.detach(), density=density.detach(), feat=feat.detach(), selector=sel, xn=xn)


def golden_mesh_finetune():
    """f-3: the reference's MeshFinetune (update_d x2, update_faces, reset_d) and the prune pass's scatter_max,
    executed over the torch_scatter stand-in."""
    import mesh_utils as MU
    from torch_scatter import scatter_max
    from oracle import quadfield_oracle as O
    verts, faces = O.shell_mesh([0.5, 0.8], subdivisions=2, jitter=1e-3, seed=11)
    g = torch.Generator().manual_seed(12)
    F, M = faces.shape[0], 5000
    mf = MU.MeshFinetune(verts.copy(), faces.astype(np.int64), 0.01)
    out = dict(verts=verts, faces=faces, scaling=np.float32(0.01))
    for it in range(2):
        d = 0.02 * torch.randn(M, 3, generator=g)
        w = torch.rand(M, generator=g)
        idx = torch.randint(0, F - 7, (M,), generator=g)          # the last 7 triangles never get a sample
        mf.update_d(d, w, idx)
        out.update({f"d{it}": d, f"w{it}": w, f"idx{it}": idx})
    out["cache_d"], out["cache_w"] = mf.cache_d.clone(), mf.cache_w.clone()
    mf.update_faces()
    out["verts_after"] = np.asarray(mf.vertices, np.float32)
    mf.reset_d()
    out["cache_w_reset"] = mf.cache_w.clone()
    # prune pass: two frames of scatter_max + maximum
    tri_w = torch.zeros(F)
    for it in range(2):
        wts = torch.rand(M, 1, generator=g) - 0.1                   # a few negative entries
        idx = torch.randint(0, F - 7, (M,), generator=g)
        cur = torch.zeros_like(tri_w)
        scatter_max(wts[:, 0], idx, out=cur)
        tri_w = torch.maximum(tri_w, cur)
        out.update({f"pw{it}": wts, f"pidx{it}": idx})
    out["tri_w"] = tri_w
    return out


def golden_field_net():
    """f-2: the reference's quadrature `Field` (train_field.py:238-252 configuration, smaller table) executed over the
    tinycudann stand-in: forward, field_grad (create_graph), compute_field_loss and loss.backward()."""
    import field as FM
    out = {}
    # small tables keep the fixture small: (a) 3 dense + 13 hashed levels, (b) all levels hashed
    for tag, hidden, nl, log2_T, min_res in (("a", 16, "elu", 10, 4), ("b", 32, "relu", 11, 16)):
        torch.manual_seed(21 + hidden)
        net = FM.Field(scale=0.5, precision=16, log2_T=log2_T, L=16, max_res=512, min_res=min_res, output_dim=1,
                       hidden_size=hidden, num_features=2, back_prop=False, nl=nl, bias=True, bias_last=True)
        g = torch.Generator().manual_seed(22)
        with torch.no_grad():                                 # features large enough to matter next to the xyz inputs
            net.xyz_encoder.params.copy_(((torch.rand(net.xyz_encoder.params.shape, generator=g) * 2 - 1) * 0.5).half().float())
        M = 800
        x = (torch.rand(M, 3, generator=g) - 0.5) * 0.98
        w, wr = torch.rand(M, generator=g), torch.rand(M, generator=g)
        dirs = torch.randn(M, 3, generator=g)
        xin = x.clone().requires_grad_(True)
        fld, fgrad = net(xin)
        loss = net.compute_field_loss(w, weights_rev=wr, field_norm=fgrad, view_dirs=dirs)
        (loss + 0.5 * fld.pow(2).mean()).backward()          # exercise the plain output path too
        sd = net.state_dict()
        for k, v in sd.items():
            out[f"{tag}_p_{k}"] = v
        for k, p in net.named_parameters():
            out[f"{tag}_g_{k}"] = p.grad
        out.update({f"{tag}_x": x, f"{tag}_w": w, f"{tag}_wr": wr, f"{tag}_dirs": dirs, f"{tag}_field": fld, f"{tag}_field_grad": fgrad,
                    f"{tag}_loss": loss, f"{tag}_hidden": np.int64(hidden), f"{tag}_nl": np.array(nl),
                    f"{tag}_log2_T": np.int64(log2_T), f"{tag}_min_res": np.int64(min_res)})
        out[f"{tag}_p_xyz_encoder.params"] = sd["xyz_encoder.params"].half()      # fp16-representable by construction
        fld2, none = net(x.clone(), return_grad=False)
        assert none is None
        out[f"{tag}_field_nograd"] = fld2
    return out


def golden_finetune_step():
    """f-3 end to end: the reference's `render_image_finetune_with_occgrid` (utils.py:465-607) in TRAINING mode —
    deformation field on the quadrature points and on random points of their triangles, re-sort, radiance field at the
    moved points, derive_properties, regulariser, MeshFinetune.update_d — followed by loss.backward(), with the
    reference Field / NGPRadianceField / MeshFinetune / MeshIntersection classes over the stand-ins."""
    import field as FM
    import mesh_utils as MU
    import radiance_fields.ngp as NGP
    import trimesh
    import utils as U
    from datasets.utils import Rays
    from oracle import quadfield_oracle as O
    verts, faces = O.shell_mesh([0.5, 0.8, 1.0], subdivisions=2, jitter=1e-3, seed=3)
    f, cx, cy, W, H = O.pinhole_intrinsics(20, 20, 0.6911)
    origins, viewdirs = O.generate_rays(O.look_at_c2w((2.4, 1.9, 1.3)), W, H, f, cx, cy)
    K, step, scaling = 6, 0.005, 1.0 / 128
    mesh = trimesh.Trimesh(vertices=verts, faces=faces, process=False)

    class FakeIntersector:
        def intersects_id(self, o, v, multiple_hits=True, return_locations=True, max_hits=10):
            return O.intersects_id(o, v, verts, faces, max_hits)

    mi = object.__new__(MU.MeshIntersection)
    mi.mesh, mi.num_intersections, mi.render_step_size = mesh, K, step
    mi.rayintersector = FakeIntersector()
    mi.vertices = torch.from_numpy(verts.astype(np.float32))
    points, vectors, index_ray, depth, index_tri, _, org = mi.sampling_raytrace_numpy(viewdirs, origins)
    data = tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in
                 (points.astype(np.float32), vectors.astype(np.float32), index_ray, depth.astype(np.float32), index_tri,
                  org.astype(np.float32)))
    log2_T = 12
    p = O.make_ngp_params(seed=42, log2_hashmap_size=log2_T, table_scale=1e4)
    p.base_w[1][0] = (p.base_w[1][0].abs() * 8.0).half().float()          # opacity spans (0,1), colours vary (scene.py)
    p.head_w[2][:3] = (p.head_w[2][:3] * 4.0).half().float()
    rf = NGP.NGPRadianceField(aabb=p.aabb.tolist(), log2_hashmap_size=log2_T)
    with torch.no_grad():
        rf.mlp_base.params.copy_(torch.cat([w.flatten() for w in p.base_w] + [p.table.flatten()]))
        rf.mlp_head.params.copy_(torch.cat([w.flatten() for w in p.head_w]))
    rf_base0, rf_head0 = rf.mlp_base.params.detach().clone(), rf.mlp_head.params.detach().clone()
    torch.manual_seed(31)
    net = FM.Field(scale=1.5, precision=16, log2_T=10, L=16, max_res=512, min_res=16, output_dim=1, hidden_size=32,
                   num_features=2, back_prop=False, nl="relu")                         # train_finetune.py:387-399
    g = torch.Generator().manual_seed(32)
    with torch.no_grad():
        net.xyz_encoder.params.copy_(((torch.rand(net.xyz_encoder.params.shape, generator=g) * 2 - 1) * 0.5).half().float())
    mf = MU.MeshFinetune(verts.copy(), faces.astype(np.int64), scaling)
    rays = Rays(origins=torch.from_numpy(origins), viewdirs=torch.from_numpy(viewdirs))
    pixels = torch.rand(origins.shape[0], 3, generator=g)
    M = data[0].shape[0]
    torch.manual_seed(33)
    bary = torch.rand((M, 3))
    torch.manual_seed(33)                                     # the function's first RNG call draws the same numbers
    prev = O.ROUND_HIDDEN
    O.ROUND_HIDDEN = True                                     # tcnn precision: fp16 hidden activations
    try:
        rgb, opacity, depth_img, n, weights, positions, index_ray_s, loss_reg, index_tri_o = U.render_image_finetune_with_occgrid(
            rf, net, None, rays, data, render_step_size=step, render_bkgd=None, mesh_intersect=mi, mesh_finetune=mf,
            scaling=scaling, bg_color="white")
        loss = torch.nn.functional.smooth_l1_loss(rgb.squeeze(), pixels) + loss_reg
        loss.sum().backward()
    finally:
        O.ROUND_HIDDEN = prev
    out = dict(verts=verts, faces=faces, origins=origins, viewdirs=viewdirs, K=np.int64(K), step=np.float32(step),
               scaling=np.float32(scaling), log2_T=np.int64(log2_T), pixels=pixels, bary=bary,
               data_xyzs=data[0], data_dirs=data[1], data_index_ray=data[2], data_ts=data[3], data_index_tri=data[4], data_origins=data[5],
               rgb=rgb, opacity=opacity, depth=depth_img, n=np.int64(n), weights=weights, positions=positions,
               index_ray=index_ray_s, loss_reg=loss_reg, index_tri=index_tri_o, cache_d=mf.cache_d, cache_w=mf.cache_w,
               g_base=rf.mlp_base.params.grad, g_head=rf.mlp_head.params.grad, rf_base=rf_base0.half(), rf_head=rf_head0.half())
    for k, v in net.state_dict().items():
        out[f"p_{k}"] = v.half() if k == "xyz_encoder.params" else v
    for k, v in net.named_parameters():
        out[f"g_{k}"] = v.grad
    return out


def golden_sg_field():
    """a8: the reference's `NGPRadianceFieldSGNew` (use_viewdirs=False, as every script builds it): `features`, `forward`
    and `features_to_rgb` over the tinycudann stand-in."""
    import radiance_fields.ngp as NGP
    from oracle import quadfield_oracle as O
    log2_T, L = 12, 3
    p = O.make_ngp_params(seed=7, log2_hashmap_size=log2_T, table_scale=1e4)
    p.base_w[1][0] = (p.base_w[1][0].abs() * 4.0).half().float()
    torch.manual_seed(41)
    rf = NGP.NGPRadianceFieldSGNew(aabb=p.aabb.tolist(), use_viewdirs=False, num_g_lobes=L, log2_hashmap_size=log2_T)
    with torch.no_grad():
        rf.mlp_base.params.copy_(torch.cat([w.flatten() for w in p.base_w] + [p.table.flatten()]))
        for prm in rf.mlp_head.parameters():                     # larger than the default init so that the lobes matter
            prm.mul_(3.0)
    g = torch.Generator().manual_seed(42)
    x = (torch.rand(500, 3, generator=g) * 2 - 1) * 1.45
    x[:8] *= 1.2
    d = torch.nn.functional.normalize(torch.randn(500, 3, generator=g), dim=-1)
    with torch.no_grad():
        feats = rf.features(x)
        rgb, density = rf(x, d)
        rgb2 = rf.features_to_rgb(feats[:, :-1], d)              # callers strip the density column
    out = dict(log2_T=np.int64(log2_T), L=np.int64(L), x=x, d=d, features=feats, rgb=rgb, density=density, rgb_from_features=rgb2)
    for k, v in rf.state_dict().items():
        out["p_" + k] = v.half() if k == "mlp_base.params" else v
    return out


def _np(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def golden_ray_generation():
    """a1: the reference's `SubjectLoader` (datasets/nerf_synthetic.py:157-378) executed on a tiny NeRF-synthetic-shaped
    dataset written to a temporary directory (transforms_*.json + RGBA PNGs, non-square so that a W/H swap shows): file
    loading, the eval branch (row-major rays of a whole image, upsample 1 and 2, white background) and the training
    branch (random images / pixels, random background colour, `add_ray_direction_noise`).  The random draws are
    replayed with the same seed and the same call order to record which (image, x, y) each training ray belongs to."""
    import json
    import tempfile
    from PIL import Image
    import datasets.nerf_synthetic as NS
    rng = np.random.RandomState(17)
    W, H, n_img = 12, 10, 3
    out = {}
    with tempfile.TemporaryDirectory() as root:
        os.makedirs(os.path.join(root, "toy", "train"))
        os.makedirs(os.path.join(root, "toy", "test"))
        for split in ("train", "test"):
            frames = []
            for i in range(n_img):
                img = rng.randint(0, 256, size=(H, W, 4)).astype(np.uint8)
                Image.fromarray(img, mode="RGBA").save(os.path.join(root, "toy", split, f"r_{i}.png"))
                ang = 0.7 * i + (0.3 if split == "test" else 0.0)
                eye = np.array([4.0 * np.cos(ang), 4.0 * np.sin(ang), 1.0 + 0.5 * i])
                fwd = -eye / np.linalg.norm(eye)
                right = np.cross(fwd, [0.0, 0.0, 1.0]); right /= np.linalg.norm(right)
                up = np.cross(right, fwd)
                m = np.eye(4)
                m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
                frames.append({"file_path": f"./{split}/r_{i}", "transform_matrix": m.tolist()})
            with open(os.path.join(root, "toy", f"transforms_{split}.json"), "w") as f:
                json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, f)
        for up in (1, 2):
            ds = NS.SubjectLoader("toy", root, "test", num_rays=None, upsample=up)
            d = ds[1]
            out[f"eval{up}_origins"], out[f"eval{up}_viewdirs"] = d["rays"].origins, d["rays"].viewdirs
            out[f"eval{up}_pixels"], out[f"eval{up}_color_bkgd"] = d["pixels"], d["color_bkgd"]
            out[f"eval{up}_K"] = ds.K
        out["images_test"], out["camtoworlds_test"], out["focal"] = ds.images, ds.camtoworlds, np.float64(ds.focal / 2)
        for tag, kw in (("train", {}), ("train_noise", {"add_ray_direction_noise": True}), ("train_single", {"batch_over_images": False})):
            n = 40
            ds = NS.SubjectLoader("toy", root, "train", color_bkgd_aug="random", num_rays=n, upsample=2, **kw)
            torch.manual_seed(123)
            d = ds[2]
            torch.manual_seed(123)                       # replay the draws of fetch_data (:293-309) in the same order
            image_id = torch.randint(0, n_img, (n,)) if kw.get("batch_over_images", True) else torch.full((n,), 2)
            x = torch.randint(0, ds.WIDTH, (n,))
            y = torch.randint(0, ds.HEIGHT, (n,))
            if kw.get("add_ray_direction_noise"):
                xf = x.float() + torch.rand_like(x.float())
                yf = y.float() + torch.rand_like(y.float())
                out[tag + "_xf"], out[tag + "_yf"] = xf, yf
            out[tag + "_color_bkgd_replayed"] = torch.rand(3)
            out[tag + "_image_id"], out[tag + "_x"], out[tag + "_y"] = image_id, x, y
            out[tag + "_origins"], out[tag + "_viewdirs"] = d["rays"].origins, d["rays"].viewdirs
            out[tag + "_pixels"], out[tag + "_color_bkgd"] = d["pixels"], d["color_bkgd"]
        out["images_train"], out["camtoworlds_train"] = ds.images, ds.camtoworlds
        out["train_K"] = ds.K
        # the dataset itself (PNG bytes + json), so the product's reader is tested on the very files the reference read
        for split in ("train", "test"):
            out[f"json_{split}"] = np.frombuffer(open(os.path.join(root, "toy", f"transforms_{split}.json"), "rb").read(), dtype=np.uint8)
            for i in range(n_img):
                out[f"png_{split}_{i}"] = np.frombuffer(open(os.path.join(root, "toy", split, f"r_{i}.png"), "rb").read(), dtype=np.uint8)
    return out


def main():
    _install()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, fn in (("field_rendering", golden_field_rendering), ("sg_decode", golden_sg_decode),
                     ("geometry", golden_geometry), ("derive_properties", golden_derive_properties),
                     ("ngp", golden_ngp), ("mesh_finetune", golden_mesh_finetune),
                     ("field_net", golden_field_net), ("finetune_step", golden_finetune_step),
                     ("sg_field", golden_sg_field), ("ray_generation", golden_ray_generation)):
        if only and name not in only:
            continue
        data = {k: _np(v) for k, v in fn().items()}
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"wrote {path}: {len(data)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
