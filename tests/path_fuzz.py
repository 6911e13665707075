"""Random RawPaths for the front-end property tests (CPU: against the reference front end through
the scene player; GPU: device build against the host build)."""
import numpy as np


def random_paths(seed, n_paths, strokes=True, width=3840, height=2160, clockwise=False):
    """Random RawPaths in the dump format: lines, cubics (generic, cusps, loops, degenerate),
    closed / open / move-only contours, random matrices and stroke styles. Returns the PathDump
    and the stroke thickness per path (0 for fills). clockwise: a third of the fills are
    FillRule::clockwise and a third of all matrices left-handed (clockwise fills flip their contour
    directions and negate their coverage there)."""
    from rive_runtime_b200 import front_end as F
    rng = np.random.default_rng(seed)
    verbs, points, paths = [], [], np.zeros(n_paths, dtype=F.PATH_DTYPE)
    nv = npnt = 0
    thickness = np.zeros(n_paths, np.float32)
    for i in range(n_paths):
        side = float(rng.choice([1.0, 30.0, 300.0, 1000.0]))
        pv, pp = [], []
        for _ in range(int(rng.integers(1, 4))):
            cur = rng.uniform(-side, side, 2).astype(np.float32)
            start = cur.copy()
            pv.append(0)
            pp.append(cur)
            for _ in range(int(rng.integers(0, 6))):
                kind = int(rng.integers(0, 8))
                end = rng.uniform(-side, side, 2).astype(np.float32)
                if kind < 2:
                    pv.append(1)
                    pp.append(end)
                else:
                    c1 = rng.uniform(-side, side, 2).astype(np.float32)
                    c2 = rng.uniform(-side, side, 2).astype(np.float32)
                    if kind == 2:      # collinear overshoot: a cusp
                        d = end - cur
                        c1, c2 = cur + d * np.float32(1.5), cur - d * np.float32(.5)
                    elif kind == 3:    # coincident control points
                        c1, c2 = cur.copy(), end.copy()
                    elif kind == 4:    # a loop
                        d = end - cur
                        perp = np.array([d[1], -d[0]], np.float32)
                        c1, c2 = end + perp, cur + perp
                    pv.append(4)
                    pp.extend([c1.astype(np.float32), c2.astype(np.float32), end])
                cur = end
            if rng.integers(0, 4) == 0 and len(pp) > 1:
                pv.append(1)
                pp.append(start)
            if rng.integers(0, 2) == 0:
                pv.append(5)
        ang = rng.uniform(-3.2, 3.2)
        sx, sy = rng.uniform(.1, 4.0, 2)
        m = np.array([np.cos(ang) * sx, np.sin(ang) * sx, -np.sin(ang) * sy, np.cos(ang) * sy,
                      rng.uniform(0, width), rng.uniform(0, height)], np.float32)
        if rng.integers(0, 3) == 0:
            m[1] = m[2] = 0
        if clockwise and rng.integers(0, 3) == 0:
            m[0], m[1] = -m[0], -m[1]  # mirror x: determinant < 0
            m = m + np.float32(0)      # (no -0: a matrix that went through Renderer::transform has none)
        color = int(rng.integers(0, 1 << 32))
        if strokes and rng.integers(0, 2) == 0:
            thickness[i] = np.float32(rng.uniform(.2, 60.0))
            radius, max_scale, psr = F.stroke_scalars(m, float(thickness[i]))
            paths[i] = (nv, len(pv), npnt, 0, m, color, 1, radius, int(rng.integers(0, 3)), int(rng.integers(0, 3)), psr, max_scale, 0)
        else:
            paths[i] = (nv, len(pv), npnt, int(rng.integers(0, 3 if clockwise else 2)), m, color, 0, 0.0, 0, 0, 0.0, 0.0, 0)
        verbs.extend(pv)
        points.extend(pp)
        nv += len(pv)
        npnt += len(pp)
    return F.PathDump(paths, np.array(verbs, np.uint8), np.array(points, np.float32).reshape(-1, 2), True), thickness


def prune_empty_segments(dump, thickness):
    """What RiveRenderPath's constructor does to a RawPath (RawPath::pruneEmptySegments,
    src/math/raw_path.cpp:353-413; renderer/src/rive_render_path.cpp:16-21): lines and cubics whose
    points all equal the preceding point are dropped. The GPU front end takes the RawPath as the
    renderer holds it, so randomly generated paths are pruned the same way before they are used."""
    from rive_runtime_b200 import front_end as F
    paths = dump.paths.copy()
    verbs, points = [], []
    nv = npnt = 0
    for i, p in enumerate(dump.paths):
        v = dump.verbs[int(p["first_verb"]):int(p["first_verb"]) + int(p["verb_count"])]
        k = int(p["first_point"])
        kept_v, kept_p = [], []
        for verb in v:
            n = 1 if verb in (0, 1) else 3 if verb == 4 else 0
            pts = dump.points[k:k + n]
            empty = False
            if verb == 1:
                empty = bool((pts[0] == dump.points[k - 1]).all())
            elif verb == 4:
                empty = bool((pts[2] == pts[1]).all() and (pts[1] == pts[0]).all() and (pts[0] == dump.points[k - 1]).all())
            if not empty:
                kept_v.append(int(verb))
                kept_p.extend(pts)
            k += n
        paths[i]["first_verb"], paths[i]["verb_count"], paths[i]["first_point"] = nv, len(kept_v), npnt
        verbs.extend(kept_v)
        points.extend(kept_p)
        nv += len(kept_v)
        npnt += len(kept_p)
    return F.PathDump(paths, np.array(verbs, np.uint8), np.array(points, np.float32).reshape(-1, 2), True), thickness
