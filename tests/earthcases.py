"""Shared cases of the ground-pass (K7) tests: scene, viewport, camera yaw, earth-map size.  c2 looks across the +-180 degree seam of
the equirectangular map; the odd viewport exercises helper pixels past the image edge; the 1000 x 500 map has odd mip sizes."""
import hashlib

import numpy as np

from skyrendering_b200.renderer import Renderer, synthetic_earth_albedo

CASES = {
    "c2_seam": ("c2", 192, 108, 0.0, (256, 128)),
    "c3_odd": ("c3", 191, 107, 2.0, (1000, 500)),
    "c1_low": ("c1", 96, 54, -1.0, (256, 128)),
    "c2_nomap": ("c2", 128, 72, 0.5, None),
}
GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "earth_digests.json")


def make_renderer(case, library):
    scene, w, h, yaw, map_dims = CASES[case]
    r = Renderer(scene, w, h, library=library)
    if yaw:
        r.scene.camera_move((0.0, 0.0, 0.0), 0.0, yaw)
    r.prime()
    r.ctx.set_earth_albedo(None if map_dims is None else synthetic_earth_albedo(*map_dims, seed=3))
    return r, w, h


def run_ground_pass(case, library, device="cpu"):
    """(depth, albedo, normal, orm, level codes) as numpy arrays after K7 on a cleared depth buffer and zeroed targets."""
    r, w, h = make_renderer(case, library)
    if device == "cpu":
        depth = np.ones((h, w), np.float32)
        targets = [np.zeros((h, w, 4), dt) for dt in (np.uint8, np.int16, np.uint16)]
        r.ground_pass(depth, *targets)
        out = [depth] + targets
    else:
        import torch
        depth = torch.ones((h, w), dtype=torch.float32, device="cuda")
        targets = [torch.zeros((h, w, 4), dtype=dt, device="cuda") for dt in (torch.uint8, torch.int16, torch.uint16)]
        r.ground_pass(depth, *targets)
        r.ctx.sync()
        out = [depth.cpu().numpy()] + [t.cpu().numpy() for t in targets]
    levels = r.ctx.earth_albedo_levels() if CASES[case][4] is not None else []
    return r, out, levels


def digests(out, levels):
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    d = {"depth": sha(out[0]), "albedo": sha(out[1]), "normal": sha(out[2]), "orm": sha(out[3])}
    if levels:
        d["albedo_levels"] = sha(np.concatenate([l.reshape(-1) for l in levels]))
    d["kept_fraction"] = float((out[0] != 1).mean())
    d["albedo_mean"] = [float(x) for x in out[1][..., :3].reshape(-1, 3).mean(0)]
    return d
