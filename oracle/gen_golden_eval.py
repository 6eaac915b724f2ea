#!/usr/bin/env python
"""Generate tests/golden/eval_auroc.npz by EXECUTING THE REFERENCE'S OWN helpers.py (test infrastructure).

    python oracle/gen_golden_eval.py       # needs /root/reference; run in the builder container only

/root/reference/helpers.py is imported unmodified; matplotlib (absent here) resolves to the recording stub in
oracle/mpl_stub, so `visualize_pairwise_similarity` runs to completion and the generator reads back the (fpr, tpr) curve it
plotted, the area printed in the legend (helpers.py:105) and the two groups handed to plt.boxplot (helpers.py:128).
`pairwise_similarity`'s own known-answer block (helpers.py:266-276) is stored as well.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('DAE_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'mpl_stub'))
sys.path.insert(0, REF)

import helpers as ref_helpers  # noqa: E402  (reference code)
from matplotlib import pyplot as plt  # noqa: E402  (the stub)
from sklearn.metrics import auc  # noqa: E402


def case(seed, n, h, classes, missing, quantise):
    rng = np.random.RandomState(seed)
    centers = rng.randn(classes, h)
    labels = rng.randint(0, classes, n)
    emb = (centers[labels] * 0.6 + rng.randn(n, h)).astype(np.float32)
    if missing:
        labels = labels.copy()
        labels[rng.rand(n) < 0.15] = -1
    sim = ref_helpers.pairwise_similarity(emb, metric='cosine').astype(np.float32)
    if quantise:   # many exactly tied scores across the two groups
        sim = (np.round(sim * 8) / 8).astype(np.float32)
    del plt.CALLS[:]
    ref_helpers.visualize_pairwise_similarity(labels, sim, plot='boxplot', title='t', save_path=None)
    plot = [c for c in plt.CALLS if c[0] == 'plot'][0]
    fpr, tpr = np.asarray(plot[1][0]), np.asarray(plot[1][1])
    box = [c for c in plt.CALLS if c[0] == 'boxplot'][0]
    rel, unrel = np.asarray(box[1][0][0]), np.asarray(box[1][0][1])
    return {'emb': emb, 'labels': labels.astype(np.int32), 'sim': sim, 'auroc': np.float64(auc(fpr, tpr)),
            'legend': np.array(plot[2]['label']), 'related_sorted': np.sort(rel), 'unrelated_sorted': np.sort(unrel)}


def main():
    out = {}
    cases = {'small_clean': (1, 60, 8, 3, False, False), 'missing_labels': (2, 90, 6, 4, True, False),
             'tied_scores': (3, 120, 5, 5, True, True), 'two_classes': (4, 150, 12, 2, False, False)}
    for name, cfg in cases.items():
        for k, v in case(*cfg).items():
            out[name + '/' + k] = v
    out['cases'] = np.array(sorted(cases))
    # helpers.py:266-276 known-answer block
    out['known/x'] = np.array([[1, 1, 0, 1], [0, 1, 0, 1], [0, 1, 1, 1]], dtype=np.float32)
    out['known/sim'] = ref_helpers.pairwise_similarity([[1, 1, 0, 1], [0, 1, 0, 1], [0, 1, 1, 1]])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'eval_auroc.npz'), **out)
    for name in cases:
        print(name, float(out[name + '/auroc']), str(out[name + '/legend']), len(out[name + '/related_sorted']),
              len(out[name + '/unrelated_sorted']))


if __name__ == '__main__':
    main()
