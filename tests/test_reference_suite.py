"""The reference's own test-suite (/root/reference/automated_test.py, shipped by oracle/build_ref.sh into the
git-ignored oracle/_ref/) executed against the drop-in: `import cc3d` resolves to cc3d_b200 (tests/refsuite.py).

Deselected, each for a stated reason:
  test_stress_upper_bound_for_binary_18  - exercises reference defect D3: the reference itself only
                                           "passes" it through an xfail-like RuntimeError path on worst-case input
Everything else runs, including test_connected_components_stack (our streaming front end returns a plain array;
the test only indexes the result) and the odd-sx binary 2D-8 cases (defect D1 shows on background pixels of
intermediate outputs only, which those tests do not assert on).
"""
import pytest

import refsuite

DESELECT = None


@pytest.mark.gpu
def test_reference_suite_against_cc3d_b200(tmp_path):
  res = refsuite.run(tmp_path, "b200", k=DESELECT)
  if res is None:
    pytest.skip("oracle/_ref/automated_test.py did not travel (built by oracle/build_ref.sh where /root/reference exists)")
  rc, passed, failed, tail = res
  assert rc == 0 and failed == 0 and passed > 1300, tail


def test_harness_against_the_reference_itself(tmp_path):
  """CPU check of the harness (fastremap stand-in, cc3d alias): a slice of the suite must pass on the reference build."""
  res = refsuite.run(tmp_path, "reference", k="2d_square or 3d_cross or periodic or binary_image_2d or largest_k or return_N")
  if res is None:
    pytest.skip("oracle/_ref/automated_test.py not present")
  rc, passed, failed, tail = res
  if "reference package unavailable" in tail:
    pytest.skip("reference python layer not present (/root/reference absent)")
  assert rc == 0 and failed == 0 and passed > 150, tail
