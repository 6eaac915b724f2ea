"""main_autoencoder_triplet.py keeps the reference's flag surface (reference main_autoencoder_triplet.py:20-74), pairs every
article with a positive and a negative like datasets/articles.py:83-128, and writes / restores the reference's cache files."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import main_autoencoder_triplet as cli  # noqa: E402
from test_cli import REFERENCE_DEFAULTS  # noqa: E402
from test_io_formats import _Dirs, _tiny_corpus, _same  # noqa: E402


def test_flag_surface():
    F = cli.build_parser().parse_args([])
    for k, v in REFERENCE_DEFAULTS.items():
        if k != 'triplet_strategy':
            assert getattr(F, k) == v, k
    with pytest.raises(SystemExit):      # the reference's triplet CLI defines no --triplet_strategy
        cli.build_parser().parse_args(['--triplet_strategy', 'batch_all'])
    F = cli.check_flags(cli.build_parser().parse_args(['--model_name', 't']))
    assert F.main_dir == 't' and F.triplet_strategy == 'none'
    with pytest.raises(AssertionError):
        cli.check_flags(cli.build_parser().parse_args(['--input_format', 'tfidf']))   # cross_entropy + tf-idf (:69-70)


def test_pairing_follows_the_reference_rule():
    labels = np.array([3, 1, 3, 2, 1, 3, 7, 1])
    pos, neg, valid = cli.pair_articles(labels, min_cate=2, rng=np.random.RandomState(0))
    assert pos.tolist() == [2, 4, 5, -1, 7, -1, -1, -1]           # the next row with the same label; the last of a label has none
    assert valid.tolist() == [True, True, True, False, True, False, False, False]   # label 2 and 7 are singletons
    for i in np.flatnonzero(valid):
        assert labels[pos[i]] == labels[i] and labels[neg[i]] != labels[i]
    _, _, v1 = cli.pair_articles(np.zeros(5), rng=np.random.RandomState(0))
    assert not v1.any()                                            # one label only: nothing to contrast with


def test_cache_round_trip(tmp_path):
    _tiny_corpus(tmp_path / 'corpus.snappy.parquet', n=90)
    F = cli.check_flags(cli.build_parser().parse_args(['--model_name', 'm', '--train_row', '40', '--validate_row', '12', '--max_features',
                                                        '80', '--data_path', str(tmp_path / 'corpus.snappy.parquet')]))
    model = _Dirs(tmp_path)
    d = cli.prepare_uci_triplets(F, model, rng=np.random.RandomState(0))
    files = set(os.listdir(model.data_dir))
    for stem in ('article_binary_count_vectorized', 'article_tfidf_vectorized'):
        for mid in ('', '_validate'):
            for suf in ('', '_pos', '_neg'):
                assert stem + mid + suf + '.npz' in files            # main_autoencoder_triplet.py:186-203
    tr = d['binary']['train']
    assert tr['org'].shape == tr['pos'].shape == tr['neg'].shape == (40, tr['org'].shape[1])
    assert d['binary']['validate']['org'].shape[0] == 12
    cat = d['articles'].label_category_publish_name.values
    # positives share the anchor's label: their bag of words comes from the same 50-word window of the tiny corpus' vocabulary
    vocab = np.array(sorted(d['count_vectorizer'].vocabulary_, key=d['count_vectorizer'].vocabulary_.get))
    word_lo = np.array([int(w[1:]) for w in vocab])
    for role, same in (('pos', True), ('neg', False)):
        lo = np.array([word_lo[tr[role][i].indices].min() // 20 for i in range(40)])
        anchor_lo = np.array([word_lo[tr['org'][i].indices].min() // 20 for i in range(40)])
        assert ((lo == anchor_lo).mean() > 0.9) == same
    assert len(cat) == 40
    r = cli.restore_uci_triplets(model)
    for name in ('binary', 'tfidf'):
        for split in ('train', 'validate'):
            for role in ('org', 'pos', 'neg'):
                assert _same(d[name][split][role], r[name][split][role]), (name, split, role)
    assert d['articles'].equals(r['articles'])


def test_synthetic_triplets_shapes():
    F = cli.check_flags(cli.build_parser().parse_args(['--model_name', 's', '--synthetic', '500', '--max_features', '400', '--seed', '1']))
    tr, va = cli.prepare_synthetic_triplets(F)
    assert tr['org'].shape == tr['pos'].shape == tr['neg'].shape == (400, 400) and va['org'].shape == (100, 400)
    overlap = np.asarray(tr['org'].multiply(tr['pos']).sum(1)).ravel() / np.asarray(tr['org'].sum(1)).ravel()
    stranger = np.asarray(tr['org'].multiply(tr['neg']).sum(1)).ravel() / np.asarray(tr['org'].sum(1)).ravel()
    assert overlap.mean() > 0.6 > stranger.mean()                   # positives keep ~70 % of the anchor's words


@pytest.mark.gpu
def test_triplet_cli_end_to_end_on_synthetic():
    model = cli.main(['--model_name', 'syn_t', '--synthetic', '1000', '--max_features', '1500', '--num_epochs', '2', '--batch_size', '100',
                      '--seed', '2', '--verbose', '--verbose_step', '1', '--encode_full', '--save_tsv'])
    assert os.path.exists(model.data_dir + 'article_encoded.npy') and os.path.exists(model.tsv_dir + 'article_encoded_validate.tsv')
    assert np.isfinite(model.train_cost_batch[0]).all() and len(model.train_cost_batch[0]) == 8
