"""Seeded input states shared by tools/make_ref_inputs.py (reference harness inputs) and the tests.

"spawn" = the reference's own spawn box (simulation_controller.cu:199-221); "wide" = blood cells spread to
and beyond the vein wall so that particle-wall collisions happen within the first steps.
"""
CASES = {
    ("cfg1", "spawn"): dict(seed=1234),
    ("cfg1", "wide"): dict(seed=1234, xz_half_width=52.0, y_range=(-30.0, -380.0)),
    ("mini3", "spawn"): dict(seed=1234, y_range=(-20.0, -60.0)),
    ("mini3", "wide"): dict(seed=4321, xz_half_width=52.0, y_range=(-30.0, -120.0)),
    ("cfg2", "spawn"): dict(seed=1234),
    ("cfg2", "wide"): dict(seed=1234, xz_half_width=50.0, y_range=(-30.0, -400.0)),
    ("cfg3", "vein"): dict(seed=1234, xz_half_width=34.0, y_range=(-30.0, -400.0)),
}
