"""CPU: the evaluation oracle (oracle/eval_oracle.py) against the outputs of the reference's own
helpers.visualize_pairwise_similarity / pairwise_similarity stored in tests/golden/eval_auroc.npz (oracle/gen_golden_eval.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import eval_oracle  # noqa: E402

GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'eval_auroc.npz'))
CASES = [str(c) for c in GOLD['cases']]


@pytest.mark.parametrize('name', CASES)
def test_groups_and_auroc_match_the_reference_run(name):
    labels, sim = GOLD[name + '/labels'], GOLD[name + '/sim']
    rel, unrel = eval_oracle.related_unrelated(labels, sim)
    assert np.array_equal(np.sort(rel), GOLD[name + '/related_sorted'])
    assert np.array_equal(np.sort(unrel), GOLD[name + '/unrelated_sorted'])
    want = float(GOLD[name + '/auroc'])
    got, twice = eval_oracle.auroc(rel, unrel)
    assert abs(got - want) < 1e-12
    assert abs(eval_oracle.auroc_sklearn(rel, unrel) - want) < 1e-12
    assert str(GOLD[name + '/legend']) == 'ROC curve (area = %0.2f)' % got   # the text the reference puts in its legend
    assert twice == int(round(want * 2 * len(rel) * len(unrel)))


def test_auroc_limits_and_ties():
    assert eval_oracle.auroc([2.0, 3.0], [0.0, 1.0])[0] == 1.0
    assert eval_oracle.auroc([0.0, 1.0], [2.0, 3.0])[0] == 0.0
    assert eval_oracle.auroc([1.0] * 5, [1.0] * 7)[0] == 0.5
    assert eval_oracle.auroc([1.0, 2.0], [1.0, 2.0]) == (0.5, 4)


def test_box_stats_against_numpy():
    d = np.random.RandomState(0).randn(1001)
    d[:3] = [-9.0, 8.0, 7.5]   # outliers beyond the whiskers
    s = eval_oracle.box_stats(d)
    assert s['n'] == 1001 and s['q1'] < s['median'] < s['q3']
    assert s['whisker_lo'] > -9.0 and s['whisker_hi'] < 7.5
    assert s['whisker_lo'] == d[d >= s['q1'] - 1.5 * (s['q3'] - s['q1'])].min()


def test_known_answer_block_of_reference_helpers():
    want = np.array([[0., 0.816496580927726, 0.6666666666666669], [0.816496580927726, 0., 0.816496580927726],
                     [0.6666666666666669, 0.816496580927726, 0.]])
    assert np.array_equal(GOLD['known/sim'], want)   # helpers.py:269-276, evaluated by the reference itself
