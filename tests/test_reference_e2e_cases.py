"""The reference's own end-to-end measure cases (test/cases/measure/data: input/*.yaml, want/*.yaml, testdata/*.json,
transcribed into tests/golden/e2e_cases.json by tests/golden/make_e2e_fixtures.py): group-by + sum/count/min/max/mean,
Top-N, a tag filter -- incl. the float MEAN 284.01366666666667 and the truncating int MEAN.  The oracle must reproduce
the expected rows; test_gpu_parity.py::test_reference_e2e_cases_on_the_device runs the same cases through the C ABI."""
import pytest

from oracle import oracle as O

from tests.helpers import E2E_CASES, check_e2e_rows, load_e2e_case


@pytest.mark.parametrize("name", E2E_CASES)
def test_oracle_reproduces_reference_case(name):
    part, oq, names, want, ordered = load_e2e_case(name)
    res = O.run_query(oq)
    check_e2e_rows(res, names, want, ordered, name)
