"""TEST INFRASTRUCTURE -- CPU restatement of the reference's related-vs-unrelated evaluation.

Only tests/ and oracle/gen_golden_eval.py import this; the product (dae_rnn_news_recommendation_b200.helpers) never does.
Parity pinned: tests/golden/eval_auroc.npz holds the outputs of the reference's own helpers.visualize_pairwise_similarity
(run by oracle/gen_golden_eval.py with matplotlib replaced by the recording stub in oracle/mpl_stub) and
tests/test_oracle_golden.py checks this restatement against them.
"""
import numpy as np


def related_unrelated(labels, sim):
    """Scores of the related (same label) and unrelated (different label) pairs of the strict lower triangle; rows whose
    label is negative are dropped (reference helpers.py:88-97)."""
    labels = np.asarray(labels).reshape(-1)
    sim = np.asarray(sim)
    assert sim.shape == (labels.shape[0], labels.shape[0])
    ok = labels >= 0
    valid = ok[:, None] & ok[None, :]
    same = (labels[:, None] == labels[None, :]) & valid
    lower = np.tril(np.ones_like(same, dtype=bool), -1)
    return sim[same & lower], sim[(~same) & valid & lower]


def auroc(related, unrelated):
    """Area under the ROC curve with 'Related' as the positive class (reference helpers.py:99-100: sklearn roc_curve + auc).
    Stated as the Mann-Whitney statistic in exact integer arithmetic: (#(r > u) + #(r == u)/2) / (R * U)."""
    r = np.sort(np.asarray(related, dtype=np.float64))
    u = np.sort(np.asarray(unrelated, dtype=np.float64))
    below = np.searchsorted(u, r, side='left').astype(np.int64)
    upto = np.searchsorted(u, r, side='right').astype(np.int64)
    twice = int(2 * below.sum() + (upto - below).sum())
    return twice / (2.0 * len(r) * len(u)), twice


def auroc_sklearn(related, unrelated):
    """The reference's literal call sequence (helpers.py:99-100)."""
    from sklearn.metrics import roc_curve, auc
    fpr, tpr, _ = roc_curve(['Related'] * len(related) + ['Unrelated'] * len(unrelated), list(related) + list(unrelated),
                            pos_label='Related')
    return float(auc(fpr, tpr))


def box_stats(data):
    """What plt.boxplot draws for one group (helpers.py:128): quartiles by linear interpolation and Tukey whiskers at the
    most extreme data within 1.5 IQR of the box."""
    d = np.sort(np.asarray(data, dtype=np.float64))
    q1, med, q3 = np.percentile(d, [25, 50, 75])
    iqr = q3 - q1
    lo = d[d >= q1 - 1.5 * iqr]
    hi = d[d <= q3 + 1.5 * iqr]
    return {'q1': q1, 'median': med, 'q3': q3, 'whisker_lo': float(lo.min()) if len(lo) else q1,
            'whisker_hi': float(hi.max()) if len(hi) else q3, 'mean': float(d.mean()), 'n': int(len(d))}
