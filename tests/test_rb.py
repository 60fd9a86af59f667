"""Rutherford-Boeing input (sylver_b200/rb.py), the matrix format of the reference's drivers."""
import numpy as np
import pytest

from sylver_b200 import gen, rb

# A 5x5 "rsa" file written by hand in the layout SPRAL's rb_write produces (header formats
# rutherford_boeing.f90:716-728), lower triangle of
#   [ 11  .  31  .   . ]
#   [  . 22   .  .  52 ]
#   [ 31  .  33  .   . ]        values in D-exponent and exponent-letter-free forms as Fortran
#   [  .  .   . 44   . ]        list-directed writers may emit them
#   [  . 52   .  .  55 ]
FIXTURE = (
    f"{'hand written test matrix':<72}{'TEST5':<8}\n"
    f"{5:14d} {1:13d} {1:13d} {3:13d}\n"
    f"{'rsa':<3}{'':11}{5:14d} {5:13d} {7:13d} {0:13d}\n"
    f"{'(40i2)':<16}{'(40i2)':<16}{'(3e24.16)':<20}\n"
    " 1 3 5 6 7 8\n"
    " 1 3 2 5 3 4 5\n"
    "  1.1000000000000000E+01  3.1000000000000000D+01  2.2000000000000000e+01\n"
    "  5.2000000000000000E+01  3.3000000000000000E+01  4.4000000000000000E+01\n"
    "  5.5000000000000000E+01\n"
)


def test_reads_a_hand_written_file(tmp_path):
    p = tmp_path / "m.rb"
    p.write_text(FIXTURE)
    m = rb.read(str(p))
    assert (m["title"], m["key"], m["type"], m["m"], m["n"]) == ("hand written test matrix", "TEST5", "rsa", 5, 5)
    assert m["ptr"].tolist() == [1, 3, 5, 6, 7, 8]
    assert m["row"].tolist() == [1, 3, 2, 5, 3, 4, 5]
    assert m["val"].tolist() == [11.0, 31.0, 22.0, 52.0, 33.0, 44.0, 55.0]


@pytest.mark.parametrize("kind,k", [("lap7", 6), ("kkt", 4)])
def test_round_trip(tmp_path, kind, k):
    n, ptr, row, val = gen.laplacian_7pt(k) if kind == "lap7" else gen.stokes_kkt(k)
    rng = np.random.default_rng(5)
    val = val * 10.0 ** rng.uniform(-200, 200, len(val))        # three-digit exponents too
    p = tmp_path / "a.rb"
    rb.write(str(p), n, ptr, row, val, title="round trip", key="RT")
    m = rb.read(str(p))
    assert m["type"] == "rsa" and m["n"] == n
    assert np.array_equal(m["ptr"], ptr[: n + 1]) and np.array_equal(m["row"], row)
    assert np.allclose(m["val"], val, rtol=1e-15, atol=0)
    head = p.read_text().splitlines()[:4]
    assert len(head[0]) == 80 and head[2].startswith("rsa") and all(len(h) <= 80 for h in head)
    # pattern only
    rb.write(str(p), n, ptr, row, None)
    m = rb.read(str(p))
    assert m["type"] == "psa" and m["val"] is None and np.array_equal(m["row"], row)


def test_rejects_what_it_cannot_read(tmp_path):
    p = tmp_path / "e.rb"
    p.write_text(FIXTURE.replace("rsa", "rse"))
    with pytest.raises(ValueError):
        rb.read(str(p))
    p.write_text(FIXTURE.replace("rsa", "csa"))
    with pytest.raises(ValueError):
        rb.read(str(p))
