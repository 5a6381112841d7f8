"""Size-independent properties of the integer side of the path (voxel grouping, neighbour tables), written with plain
torch ops on whatever device the tensors live on.  They are what the parity suite asserts at sizes the CPU oracle cannot
reach in seconds (BASELINE config 5: 100 k-point clouds at voxel 0.025 m, 64-bit cell keys); tests/test_properties_cpu.py
pins the checkers themselves against the oracle's tables (and against deliberately corrupted ones) at small sizes."""
import torch


def flat_keys(coords, shape):
    """(b, z, y, x) rows -> int64 cell keys in spconv's flat order."""
    c = coords.long()
    d, h, w = (int(s) for s in shape)
    return ((c[:, 0] * d + c[:, 1]) * h + c[:, 2]) * w + c[:, 3]


def tap_offsets(ksize, device):
    """[K, 3] (kz, ky, kx) of every kernel tap, row-major, last fastest."""
    kz, ky, kx = (int(k) for k in ksize)
    g = torch.stack(torch.meshgrid(torch.arange(kz), torch.arange(ky), torch.arange(kx), indexing="ij"), -1)
    return g.reshape(-1, 3).to(device)


def check_voxelization(points, scene_offsets, voxels, coords, num_points, n_voxels, voxel_size, coors_range, grid,
                       max_points, max_voxels):
    """Voxel rows are distinct cells, counts respect the caps, every stored point lies in its voxel's cell, padding
    slots are zero, the per-scene counts add up, and (when no cap binds) every in-range point was stored exactly once."""
    dev = points.device
    B = scene_offsets.numel() - 1
    nv = n_voxels.cpu().tolist()
    m = nv[B]
    assert sum(nv[:B]) == m and all(0 <= v <= max_voxels for v in nv[:B])
    coords, voxels, num_points = coords[:m], voxels[:m], num_points[:m]
    lo = torch.tensor(coors_range[:3], dtype=torch.float32, device=dev)
    vs = torch.tensor(voxel_size, dtype=torch.float32, device=dev)
    gx, gy, gz = (int(g) for g in grid)
    keys = flat_keys(coords, [gz, gy, gx])
    assert torch.unique(keys).numel() == m, "two voxel rows share a cell"
    assert int(num_points.min()) >= 1 and int(num_points.max()) <= max_points
    # scene index column: rows of scene b are contiguous and in scene order
    assert bool((coords[1:, 0] >= coords[:-1, 0]).all())
    assert torch.equal(torch.bincount(coords[:, 0].long(), minlength=B).cpu(), torch.tensor(nv[:B]))
    # every stored point quantises (fp32 subtract, IEEE divide, floor) to its voxel's (z, y, x)
    P = voxels.shape[1]
    slot = torch.arange(P, device=dev).view(1, P)
    live = slot < num_points.view(-1, 1)
    q = torch.floor((voxels[..., :3] - lo) / vs).to(torch.int32)            # (x, y, z) order
    want = coords[:, [3, 2, 1]].view(m, 1, 3).expand(-1, P, -1)
    assert torch.equal(q[live], want[live]), "a stored point lies outside its voxel"
    assert float(voxels[~live].abs().sum()) == 0.0, "padding slots are not zero"
    # conservation when no cap binds
    if all(v < max_voxels for v in nv[:B]) and int(num_points.max()) < max_points:
        n_total = int(scene_offsets[B].item())
        p = points[:n_total, :3]
        f = torch.floor((p - lo) / vs)
        g = torch.tensor([gx, gy, gz], dtype=torch.float32, device=dev)
        in_range = ((f >= 0) & (f < g)).all(1)
        assert int(num_points.sum()) == int(in_range.sum()), "points lost or duplicated"
    return m


def check_subm_table(coords, nbr, shape, ksize, dilation=1):
    """Sub-manifold neighbour table [N, K]: centre tap = the row itself, every entry points at the row whose coordinate is
    the shifted coordinate, the relation is its own mirror image, and no existing neighbour is missing."""
    dev = coords.device
    n, K = nbr.shape
    ks = [int(k) for k in ksize]
    off = (tap_offsets(ks, dev) - torch.tensor([k // 2 for k in ks], device=dev)) * int(dilation)
    rows = torch.arange(n, device=dev)
    assert torch.equal(nbr[:, K // 2].long(), rows), "centre tap is not the site itself"
    i, k = torch.nonzero(nbr >= 0, as_tuple=True)
    j = nbr[i, k].long()
    assert int(j.max()) < n
    assert torch.equal(coords[j, 0], coords[i, 0]), "neighbour in another scene"
    assert torch.equal(coords[j, 1:].long(), coords[i, 1:].long() + off[k]), "entry does not sit at the shifted coordinate"
    assert torch.equal(nbr[j, K - 1 - k].long(), i), "relation is not symmetric"
    # completeness: count the (site, tap) pairs whose shifted coordinate is an active site
    keys = flat_keys(coords, shape)
    dims = torch.tensor([int(s) for s in shape], device=dev)
    total = 0
    for t in range(K):
        c = coords[:, 1:].long() + off[t]
        ok = ((c >= 0) & (c < dims)).all(1)
        shifted = torch.cat([coords[ok, :1].long(), c[ok]], 1)
        total += int(torch.isin(flat_keys(shifted, shape), keys).sum())
    assert total == i.numel(), ("missing or spurious neighbours", total, i.numel())
    return i.numel()


def check_conv_tables(coords_in, out_coords, nbr_out, nbr_in, in_shape, out_shape, ksize, stride, padding, dilation=1):
    """Strided-conv rulebook: outputs are the distinct touched cells in ascending flat-key order (spconv's order), every
    table entry satisfies out * stride - pad + k * dil == in, nbr_in is the transpose of nbr_out, every output row has at
    least one input and the pair count equals the number of (input, tap) candidates that land inside the output grid."""
    dev = coords_in.device
    ks, st, pd = [int(k) for k in ksize], [int(s) for s in stride], [int(p) for p in padding]
    K = ks[0] * ks[1] * ks[2]
    assert nbr_out.shape[1] == K and nbr_in.shape[1] == K
    keys = flat_keys(out_coords, out_shape)
    assert bool((keys[1:] > keys[:-1]).all()), "outputs are not in ascending flat-key order / not distinct"
    taps = tap_offsets(ks, dev) * int(dilation)
    stv, pdv = torch.tensor(st, device=dev), torch.tensor(pd, device=dev)
    o, k = torch.nonzero(nbr_out >= 0, as_tuple=True)
    i = nbr_out[o, k].long()
    assert int(i.max()) < coords_in.shape[0]
    assert torch.equal(coords_in[i, 0], out_coords[o, 0])
    assert torch.equal(coords_in[i, 1:].long(), out_coords[o, 1:].long() * stv - pdv + taps[k]), "pair violates the geometry"
    assert torch.equal(nbr_in[i, k].long(), o), "nbr_in is not the transpose of nbr_out"
    assert int((nbr_in >= 0).sum()) == o.numel()
    assert bool((nbr_out >= 0).any(1).all()), "an output site without input"
    # candidates: for every input and tap, t = in + pad - k*dil must be divisible by the stride and land in the grid
    dims = torch.tensor([int(s) for s in out_shape], device=dev)
    total = 0
    cand_keys = []
    for t in range(K):
        num = coords_in[:, 1:].long() + pdv - taps[t]
        ok = ((num >= 0) & (num % stv == 0)).all(1)
        oc = num // stv
        ok &= (oc < dims).all(1)
        total += int(ok.sum())
        cand_keys.append(flat_keys(torch.cat([coords_in[ok, :1].long(), oc[ok]], 1), out_shape))
    assert total == o.numel(), ("pair count differs from the candidate count", total, o.numel())
    assert torch.equal(torch.unique(torch.cat(cand_keys)), keys), "output site set differs from the touched cells"
    return o.numel()
