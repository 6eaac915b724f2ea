"""main_autoencoder.py keeps the reference's flag names, defaults and asserts (reference main_autoencoder.py:27-111)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import main_autoencoder as cli  # noqa: E402

REFERENCE_DEFAULTS = dict(verbose=False, verbose_step=5, encode_full=False, validation=False, input_format='binary',
                          label='category_publish_name', save_tsv=False, train_row=8000, validate_row=2000,
                          restore_previous_data=False, min_df=0, max_df=0.99, max_features=10000, model_name='',
                          restore_previous_model=False, seed=-1, compress_factor=20, corr_type='masking', corr_frac=0.3,
                          xavier_init=1, enc_act_func='sigmoid', dec_act_func='sigmoid', main_dir='', loss_func='cross_entropy',
                          opt='gradient_descent', learning_rate=0.1, momentum=0.5, num_epochs=50, batch_size=0.1, alpha=1,
                          triplet_strategy='batch_all')


def test_defaults_match_reference():
    F = cli.build_parser().parse_args([])
    for k, v in REFERENCE_DEFAULTS.items():
        assert getattr(F, k) == v, k


def test_asserts_and_env_override(monkeypatch):
    F = cli.check_flags(cli.build_parser().parse_args(['--model_name', 'uci', '--opt', 'adam']))
    assert F.main_dir == 'uci'
    with pytest.raises(AssertionError):  # tf-idf input forbids cross-entropy (main_autoencoder.py:108-109)
        cli.check_flags(cli.build_parser().parse_args(['--input_format', 'tfidf']))
    with pytest.raises(AssertionError):
        cli.check_flags(cli.build_parser().parse_args(['--triplet_strategy', 'semi_hard']))
    monkeypatch.setenv('corr_type', 'decay')
    monkeypatch.setenv('corr_frac', '0.5')
    monkeypatch.setenv('alpha', '0.5')                    # float flags stay floats (the reference casts with float(), :90)
    monkeypatch.setenv('restore_previous_model', '0')     # an explicit false value is False, not "the variable exists"
    monkeypatch.setenv('label', 'story')                  # not in the reference's override list (:75-92): ignored
    monkeypatch.setenv('verbose', '1')
    F = cli.apply_env_overrides(cli.build_parser().parse_args([]))
    assert F.corr_type == 'decay' and F.corr_frac == 0.5  # the reference reads compress_factor here (:79-80)
    assert F.alpha == 0.5 and F.restore_previous_model is False and F.label == 'category_publish_name' and F.verbose is False
    monkeypatch.setenv('restore_previous_model', 'True')
    assert cli.apply_env_overrides(cli.build_parser().parse_args([])).restore_previous_model is True


@pytest.mark.gpu
def test_cli_end_to_end_on_synthetic():
    model = cli.main(['--model_name', 'syn', '--synthetic', '1200', '--max_features', '2000', '--num_epochs', '2', '--batch_size',
                      '200', '--seed', '3', '--verbose', '--verbose_step', '1', '--encode_full', '--validation'])
    assert os.path.exists(model.data_dir + 'article_encoded.npy') and os.path.exists(model.model_path + '.npz')
    assert os.path.exists(model.parameter_file)
    assert model.train_cost_batch[0][-1] < model.history[0][0, 0]
    ev = model.evaluation   # the reference's evaluation tail (main_autoencoder.py:307-360) as numbers
    for key in ('similarity_boxplot_binary_count(Category)', 'similarity_boxplot_encoded(Category)',
                'similarity_boxplot_binary_count_validate(Category)', 'similarity_boxplot_encoded_validate(Category)'):
        assert 0.0 <= ev[key]['auroc'] <= 1.0 and os.path.exists(model.plot_dir + key + '.json')
    idx, score = ev['nearest']
    assert idx.shape == (960,) and (idx != range(960)).all() and (score <= 1.0 + 1e-5).all()


@pytest.mark.gpu
def test_cli_zero_epochs_restores_and_encodes():
    """--num_epochs 0 "will not train the model" (reference main_autoencoder.py:71): with --restore_previous_model it is the
    restore-then-encode path; train_time stays None and the run must still save the embeddings."""
    import numpy as np
    common = ['--model_name', 'syn0', '--synthetic', '600', '--max_features', '1500', '--batch_size', '100', '--seed', '3',
              '--triplet_strategy', 'none']
    m1 = cli.main(common + ['--num_epochs', '1', '--encode_full'])
    e1 = np.load(m1.data_dir + 'article_encoded.npy')
    m2 = cli.main(common + ['--num_epochs', '0', '--restore_previous_model', '--encode_full'])
    assert m2.train_time is None
    e2 = np.load(m2.data_dir + 'article_encoded.npy')
    assert np.array_equal(e1, e2)
