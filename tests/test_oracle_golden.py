"""Pins the CPU oracle against the reference's own golden vectors."""
import pytest

import golden_runner as R
from golden import reference_vectors as G


@pytest.mark.parametrize("case", G.ENCODE_CASES, ids=[c["id"] for c in G.ENCODE_CASES])
def test_encode_golden(oracle, case):
    R.run_encode_case(oracle, case)


@pytest.mark.parametrize("case", G.DECODE_CASES, ids=[c[0] for c in G.DECODE_CASES])
def test_decode_golden(oracle, case):
    R.run_decode_case(oracle, case)


def test_seal(oracle):
    R.run_seal_case(oracle)


def test_lengths(oracle):
    R.run_length_cases(oracle)


def test_misc(oracle):
    R.run_misc_cases(oracle)
