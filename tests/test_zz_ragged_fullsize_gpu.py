"""BASELINE-size MixedOP with the ragged widths an elastic-width search leaves behind (train_search.py:293-305).
Kept in its own, last-sorted module: it was added after the round's GPU budget was spent and has not run on a B200 yet."""
import pytest

from tests import helpers as H
from tests.test_fullsize_gpu import _fullsize

pytestmark = pytest.mark.gpu


def test_fullsize_ragged_widths():
    """Odd mid widths at N=128 go through the straight-line (VEC) producers with clamped row indices and channel
    offsets that are not multiples of 4."""
    _fullsize(40, 40, 1, 'swish', 28, H.default_mcs(40, ragged=True))
