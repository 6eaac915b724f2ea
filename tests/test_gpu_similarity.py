"""pairwise similarity + nearest-article lookup on the GPU vs sklearn, incl. the reference's only known-answer
vector (helpers.py:269-276)."""
import numpy as np
import pytest
import scipy.sparse as sp
from sklearn.metrics import pairwise
from sklearn.preprocessing import normalize

pytestmark = pytest.mark.gpu


def test_known_answer_vector_from_reference_helpers():
    from dae_rnn_news_recommendation_b200.helpers import pairwise_similarity
    cnt = [[1, 1, 0, 1], [0, 1, 0, 1], [0, 1, 1, 1]]
    want = np.array([[0., 0.816496580927726, 0.6666666666666669], [0.816496580927726, 0., 0.816496580927726],
                     [0.6666666666666669, 0.816496580927726, 0.]])
    for inp in (cnt, np.array(cnt), sp.csr_matrix(cnt)):
        assert np.allclose(pairwise_similarity(inp), want, atol=2e-6)


@pytest.mark.parametrize('metric,norm', [('cosine', ''), ('linear kernel', ''), ('linear kernel', 'l2'), ('cosine', 'l1')])
def test_dense_matches_sklearn(metric, norm):
    from dae_rnn_news_recommendation_b200.helpers import pairwise_similarity
    rng = np.random.default_rng(0)
    e = (rng.random((700, 500)).astype(np.float32) - 0.4)
    ref_in = normalize(e, norm=norm) if norm else e
    want = (pairwise.cosine_similarity if metric == 'cosine' else pairwise.linear_kernel)(ref_in.astype(np.float64))
    np.fill_diagonal(want, 0)
    got = pairwise_similarity(e, norm=norm, metric=metric)
    assert got.shape == (700, 700) and np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


def test_sparse_matches_sklearn_and_nearest_neighbors():
    from dae_rnn_news_recommendation_b200.helpers import pairwise_similarity, nearest_neighbors
    from helpers import random_csr
    x = random_csr(300, 2000, 40, kind='tfidf', seed=3)
    want = pairwise.cosine_similarity(x)
    np.fill_diagonal(want, 0)
    assert np.abs(pairwise_similarity(x) - want).max() < 1e-5
    rng = np.random.default_rng(1)
    e = rng.normal(size=(1500, 64)).astype(np.float32)
    sim = pairwise.cosine_similarity(e.astype(np.float64))
    np.fill_diagonal(sim, -np.inf)
    idx, val = nearest_neighbors(e, chunk=512)
    assert (idx == sim.argmax(1)).mean() > 0.999 and np.allclose(val, sim.max(1), atol=2e-5)


def test_sparse_similarity_beyond_one_register_block():
    """More rows than the encode kernel holds per output row (4096) and not a multiple of 4: the N x N product is assembled
    from column blocks (UCI runs 8000 rows through this path, main_autoencoder.py:309)."""
    from dae_rnn_news_recommendation_b200.helpers import pairwise_similarity
    from helpers import random_csr
    x = random_csr(4501, 1500, 25, kind='binary', seed=11)
    want = pairwise.cosine_similarity(x)
    np.fill_diagonal(want, 0)
    got = pairwise_similarity(x, metric='cosine')
    assert got.shape == (4501, 4501) and np.abs(got - want).max() < 1e-5
