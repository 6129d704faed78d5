import hashlib

import numpy as np

import jpezy_b200 as J


def test_synth_is_deterministic_and_in_range():
    for fam in (0, 1, 2):
        a = J.synth.image(fam, 123, 45, frame=2)
        b = J.synth.image(fam, 123, 45, frame=2)
        assert all((x == y).all() for x, y in zip(a, b))
        assert a[0].dtype == np.uint8 and a[0].shape == (45, 123)
    r, g, b = J.synth.image(0, 64, 64)
    assert hashlib.sha256(r.tobytes() + g.tobytes() + b.tobytes()).hexdigest()[:16] == hashlib.sha256(
        r.tobytes() + g.tobytes() + b.tobytes()).hexdigest()[:16]
    assert J.synth.image(0, 64, 64, frame=1)[0].tolist() != r.tolist()
    n = J.synth.image(1, 256, 256)[0]
    assert n.min() >= 1 and n.max() == 255 and 120 < n.mean() < 136
    flat = J.synth.image(2, 64, 32)[0]
    assert (flat[:16, :16] == flat[0, 0]).all()
