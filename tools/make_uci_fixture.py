"""Vectorise the UCI news corpus exactly as the CLI does for BASELINE.json configs[0] (8000 train + 2000 validate articles,
10 000 features, binary counts; reference main_autoencoder.py:177-238) and store the result as a small fixture, so that the GPU box
(which has neither the corpus nor /root/reference) can train and evaluate on the real data.  Builder container only:

    python tools/make_uci_fixture.py [/root/reference/datasets/uci_news.snappy.parquet]   ->  tests/golden/uci_c1.npz
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import main_autoencoder as cli

path = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/datasets/uci_news.snappy.parquet'
F = cli.check_flags(cli.build_parser().parse_args(['--model_name', 'uci', '--data_path', path]))
import tempfile, types
tmp = tempfile.mkdtemp()
os.makedirs(tmp + '/data/')
d = cli.prepare_uci(F, types.SimpleNamespace(data_dir=tmp + '/data/'))     # also writes the raw-count matrices
from dae_rnn_news_recommendation_b200.io_formats import read_file
counts = (read_file(tmp + '/data/article_count_vectorized.npz'), read_file(tmp + '/data/article_count_vectorized_validate.npz'))
out = {}
for split, k in (('train', 0), ('validate', 1)):
    X = d['binary'][k].tocsr(); X.sort_indices()
    T = counts[k].tocsr(); T.sort_indices()
    assert (X.indices == T.indices).all() and (X.indptr == T.indptr).all() and T.data.max() <= 255
    out[split + '_indptr'] = X.indptr.astype(np.int32)
    out[split + '_indices'] = X.indices.astype(np.uint16)          # 10 000 features
    out[split + '_counts'] = T.data.astype(np.uint8)               # raw term counts: tf-idf is TfidfTransformer().fit(train counts)
    out[split + '_shape'] = np.array(X.shape)
    for lab in ('category_publish_name', 'story'):
        out['%s_label_%s' % (split, lab)] = np.asarray(d['label_' + lab][k]).astype(np.int32)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'uci_c1.npz'), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
