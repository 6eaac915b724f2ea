"""related-vs-unrelated AUROC + box statistics on the GPU (dae_pair_partition / dae_auroc_count through the
C ABI) against the golden outputs of the reference's own helpers.visualize_pairwise_similarity and against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'eval_auroc.npz'))
CASES = [str(c) for c in GOLD['cases']]


@pytest.mark.parametrize('name', CASES)
def test_golden_cases_of_the_reference(name):
    from dae_rnn_news_recommendation_b200.helpers import visualize_pairwise_similarity, related_unrelated_scores
    labels, sim = GOLD[name + '/labels'], GOLD[name + '/sim']
    rel, unrel = related_unrelated_scores(labels, sim)
    assert np.array_equal(rel.cpu().numpy(), GOLD[name + '/related_sorted'])      # bit-exact groups
    assert np.array_equal(unrel.cpu().numpy(), GOLD[name + '/unrelated_sorted'])
    out = visualize_pairwise_similarity(labels, sim, plot='boxplot', title=name)
    want = float(GOLD[name + '/auroc'])
    assert abs(out['auroc'] - want) < 1e-12                                        # integer count / (2RU) vs sklearn's fp64 trapezoid
    assert out['twice_u'] == int(round(want * 2 * len(GOLD[name + '/related_sorted']) * len(GOLD[name + '/unrelated_sorted'])))
    assert 'ROC curve (area = %0.2f)' % out['auroc'] == str(GOLD[name + '/legend'])


def _random_case(seed, n, classes, missing, quantise, skew=None):
    rng = np.random.RandomState(seed)
    p = None if skew is None else np.array([skew] + [(1 - skew) / (classes - 1)] * (classes - 1))
    labels = rng.choice(classes, n, p=p).astype(np.float64)   # float labels, like a pandas column
    emb = rng.randn(classes, 16)[labels.astype(int)] * 0.5 + rng.randn(n, 16)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    sim = (emb @ emb.T).astype(np.float32)
    np.fill_diagonal(sim, 0)
    if quantise:
        sim = (np.round(sim * 16) / 16).astype(np.float32)
    if missing:
        labels[rng.rand(n) < missing] = -1
    return labels, sim


@pytest.mark.parametrize('seed,n,classes,missing,quantise,skew', [
    (0, 1500, 4, 0.0, False, None),      # related < unrelated: related scores query the sorted unrelated group
    (1, 1200, 3, 0.1, True, 0.9),        # one dominant class: related > unrelated -> the swapped query direction; heavy ties
    (2, 700, 300, 0.2, True, None),      # story-like labels: very few related pairs
    (3, 33, 2, 0.0, False, None),        # rows shorter than one warp
])
def test_matches_oracle_exactly(seed, n, classes, missing, quantise, skew):
    from dae_rnn_news_recommendation_b200.helpers import visualize_pairwise_similarity
    from oracle import eval_oracle
    labels, sim = _random_case(seed, n, classes, missing, quantise, skew)
    rel, unrel = eval_oracle.related_unrelated(labels, sim)
    want, twice = eval_oracle.auroc(rel, unrel)
    out = visualize_pairwise_similarity(labels, sim)
    assert out['twice_u'] == twice and out['auroc'] == want                        # integer arithmetic: exact
    assert out['related']['n'] == len(rel) and out['unrelated']['n'] == len(unrel)
    for grp, data in (('related', rel), ('unrelated', unrel)):
        ws = eval_oracle.box_stats(data)
        for k in ('q1', 'median', 'q3', 'whisker_lo', 'whisker_hi', 'mean'):
            assert abs(out[grp][k] - ws[k]) <= 1e-6, (grp, k, out[grp][k], ws[k])  # fp32 data, fp64 interpolation on both sides


def test_degenerate_inputs():
    from dae_rnn_news_recommendation_b200.helpers import visualize_pairwise_similarity
    sim = np.zeros((5, 5), dtype=np.float32)
    out = visualize_pairwise_similarity(np.array([1, 1, 1, 1, 1]), sim)            # no unrelated pair
    assert np.isnan(out['auroc']) and out['related']['n'] == 10 and out['unrelated']['n'] == 0
    out = visualize_pairwise_similarity(np.array([-1, -1, 0, -1, 1]), sim)         # a single valid pair, tied at 0
    assert out['unrelated']['n'] == 1 and out['related']['n'] == 0 and np.isnan(out['auroc'])
    out = visualize_pairwise_similarity(np.array([0, 1, 0, 1, -1]), sim)           # every score tied
    assert out['auroc'] == 0.5


def test_full_size_complement_property_and_json(tmp_path):
    """UCI-sized evaluation (8000 articles -> 32 M pairs) straight from the device similarity matrix: AUROC(S) + AUROC(-S) = 1
    exactly (as integers: twice_u + twice_u' = 2RU), the groups hold every valid pair once, and the JSON lands next to save_path."""
    import json
    import torch
    from dae_rnn_news_recommendation_b200.helpers import pairwise_similarity, visualize_pairwise_similarity
    rng = np.random.RandomState(5)
    n, h = 8000, 500
    labels = rng.randint(0, 4, n)
    emb = (rng.randn(4, h)[labels] * 0.15 + rng.randn(n, h)).astype(np.float32)
    sim = pairwise_similarity(emb, metric='cosine', to_host=False)
    assert isinstance(sim, torch.Tensor) and sim.is_cuda
    a = visualize_pairwise_similarity(labels, sim, save_path=str(tmp_path / 'similarity_boxplot_encoded.png'))
    b = visualize_pairwise_similarity(labels, -sim)
    r, u = a['related']['n'], a['unrelated']['n']
    assert r + u == n * (n - 1) // 2
    assert a['twice_u'] + b['twice_u'] == 2 * r * u
    assert 0.5 < a['auroc'] < 1.0 and a['related']['median'] > a['unrelated']['median']
    assert abs(a['related']['median'] + b['related']['median']) < 1e-7
    saved = json.load(open(tmp_path / 'similarity_boxplot_encoded.json'))
    assert saved['auroc'] == a['auroc']
