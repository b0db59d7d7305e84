"""Shared comparison helpers for the parity tests (engine or host emulation vs the oracle)."""
import numpy as np


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a


def compare_states(sa, sb, n_arenas, tol=0.0, what=("body", "sense_aabb", "fat_aabb", "contacts", "arena", "robot")):
    """Return a list of human-readable mismatches between two abi.State (empty = identical)."""
    out = []
    for a in range(n_arenas):
        na, nb = int(sa.arena[a]["n_contacts"]), int(sb.arena[a]["n_contacts"])
        if "contacts" in what:
            if na != nb:
                out.append("arena %d: n_contacts %d vs %d" % (a, na, nb))
            n = min(na, nb)
            if not np.array_equal(sa.contact_key[a, :n], sb.contact_key[a, :n]):
                i = int(np.nonzero(sa.contact_key[a, :n] != sb.contact_key[a, :n])[0][0])
                out.append("arena %d: contact_key differs first at %d: %08x vs %08x" % (a, i, sa.contact_key[a, i], sb.contact_key[a, i]))
            elif not np.array_equal(sa.contact_info[a, :n], sb.contact_info[a, :n]):
                i = int(np.nonzero(sa.contact_info[a, :n] != sb.contact_info[a, :n])[0][0])
                out.append("arena %d: contact_info differs first at %d (key %08x): %06x vs %06x" % (a, i, sa.contact_key[a, i], sa.contact_info[a, i], sb.contact_info[a, i]))
            else:
                x, y = sa.contact_impulse[a, :n], sb.contact_impulse[a, :n]
                bad = ~close(x, y, tol)
                if bad.any():
                    i = int(np.nonzero(bad.any(axis=1))[0][0])
                    out.append("arena %d: contact_impulse differs at %d (key %08x): %s vs %s" % (a, i, sa.contact_key[a, i], x[i], y[i]))
        for name in ("body", "sense_aabb", "fat_aabb"):
            if name in what:
                x, y = getattr(sa, name)[a], getattr(sb, name)[a]
                bad = ~close(x, y, tol)
                if bad.any():
                    idx = np.argwhere(bad)[0]
                    out.append("arena %d: %s differs at %s (%d entries): %r vs %r" % (a, name, tuple(idx), int(bad.sum()), x[tuple(idx)], y[tuple(idx)]))
        if "arena" in what:
            for f in ("rng_seed", "rng_have_gaussian", "update_count", "step_count", "inv_dt0", "rng_next_gaussian"):
                x, y = sa.arena[a][f], sb.arena[a][f]
                # the cached second Gaussian comes out of a double log / sqrt: device libm vs glibc may differ in the last place
                same = x == y or (f == "rng_next_gaussian" and tol > 0 and abs(float(x) - float(y)) <= 4e-16 * max(1.0, abs(float(y))))
                if not same:
                    out.append("arena %d: arena.%s %r vs %r" % (a, f, x, y))
        if "robot" in what:
            for r in range(sa.robot.shape[1]):
                ra, rb = sa.robot[a, r], sb.robot[a, r]
                for f in ra.dtype.names:
                    if f == "pad_":
                        continue
                    x, y = np.asarray(ra[f]), np.asarray(rb[f])
                    if f == "rng_next_gaussian":   # per-robot stream (PS_RNG_PER_ROBOT): as the arena's cached Gaussian above
                        if not (x == y or (tol > 0 and abs(float(x) - float(y)) <= 4e-16 * max(1.0, abs(float(y))))):
                            out.append("arena %d robot %d: %s %r vs %r" % (a, r, f, x, y))
                        continue
                    if x.dtype.kind == "f":
                        ok = close(x, y, tol).all()
                    else:
                        ok = np.array_equal(x, y)
                    if not ok:
                        out.append("arena %d robot %d: %s %r vs %r" % (a, r, f, x, y))
    return out


def close(x, y, tol):
    x = np.asarray(x)
    y = np.asarray(y)
    if tol == 0.0:
        return (bits(x) == bits(y)) | ((x == 0) & (y == 0))
    return np.abs(x.astype(np.float64) - y.astype(np.float64)) <= tol * np.maximum(1.0, np.maximum(np.abs(x), np.abs(y)))


def compare_observations(oa, ob, tol=0.0):
    """Mismatches between two abi.Observations (camera, occupancy, local map, pose)."""
    out = []
    if oa.camera is not None and ob.camera is not None and not np.array_equal(oa.camera, ob.camera):
        idx = np.argwhere(oa.camera != ob.camera)
        out.append("camera differs at %d pixels, first (arena, robot, y, x) = %s: %d vs %d" % (
            len(idx), tuple(idx[0]), oa.camera[tuple(idx[0])], ob.camera[tuple(idx[0])]))
    if not np.array_equal(oa.occupancy, ob.occupancy):
        idx = np.argwhere(oa.occupancy != ob.occupancy)
        out.append("occupancy differs at %d cells, first %s: %d vs %d" % (len(idx), tuple(idx[0]), oa.occupancy[tuple(idx[0])], ob.occupancy[tuple(idx[0])]))
    if not close(oa.pose, ob.pose, tol).all():
        out.append("pose differs")
    A, R = oa.localmap.shape
    for a in range(A):
        for r in range(R):
            la, lb = oa.localmap[a, r], ob.localmap[a, r]
            for f in ("carrying", "carried_type", "carried_cluster", "n_pucks", "n_clusters", "left_sum", "right_sum", "overflow"):
                if la[f] != lb[f]:
                    out.append("arena %d robot %d: localmap.%s %r vs %r" % (a, r, f, la[f], lb[f]))
            for f in ("carried_x", "carried_y"):
                if not close(la[f], lb[f], tol):
                    out.append("arena %d robot %d: localmap.%s %r vs %r" % (a, r, f, la[f], lb[f]))
            if la["n_pucks"] == lb["n_pucks"]:
                n = int(la["n_pucks"])
                for f in ("puck_x", "puck_y"):
                    if not close(la[f][:n], lb[f][:n], tol).all():
                        out.append("arena %d robot %d: localmap.%s differs" % (a, r, f))
                for f in ("puck_type", "puck_cluster", "puck_degree"):
                    if not np.array_equal(la[f][:n], lb[f][:n]):
                        out.append("arena %d robot %d: localmap.%s %s vs %s" % (a, r, f, la[f][:n], lb[f][:n]))
            if la["n_clusters"] == lb["n_clusters"]:
                n = int(la["n_clusters"])
                for f in ("cluster_x", "cluster_y"):
                    if not close(la[f][:n], lb[f][:n], tol).all():
                        out.append("arena %d robot %d: localmap.%s differs" % (a, r, f))
                for f in ("cluster_type", "cluster_size", "cluster_kept"):
                    if not np.array_equal(la[f][:n], lb[f][:n]):
                        out.append("arena %d robot %d: localmap.%s %s vs %s" % (a, r, f, la[f][:n], lb[f][:n]))
    return out
