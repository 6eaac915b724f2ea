#!/usr/bin/env python
"""CLI with the reference's flag surface (reference main_autoencoder.py:23-111) driving the B200 DenoisingAutoencoder.

    python main_autoencoder.py --model_name uci --verbose --encode_full [--data_path datasets/uci_news.snappy.parquet]
    python main_autoencoder.py --model_name syn --synthetic 100000 --num_epochs 2 --batch_size 800 --verbose

Same flag names, defaults, asserts and `.env` override as the reference (python-dotenv; the two env typos at
main_autoencoder.py:79-80 are fixed and `adam` is accepted).  Data preparation follows
main_autoencoder.py:177-238 (CountVectorizer -> binary / tf-idf CSR, factorised labels).  The evaluation tail (:307-360) runs on
the GPU as numbers, not pictures: pairwise similarity of the inputs and of the embeddings, related-vs-unrelated AUROC + box
statistics (one JSON per reference plot file name under <plot_dir>) and the most-similar-article lookup; drawing with matplotlib
is not reproduced.
"""
import argparse
import os
from pathlib import Path

import numpy as np

_script_path = Path(os.path.dirname(os.path.realpath(__file__)))


def _bool_flag(ap, name, default, help_):
    ap.add_argument('--' + name, dest=name, action='store_true', default=default, help=help_)
    ap.add_argument('--no' + name, dest=name, action='store_false')


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    _bool_flag(ap, 'verbose', False, 'Level of verbosity. 0 - silent, 1 - print log')
    ap.add_argument('--verbose_step', type=int, default=5)
    _bool_flag(ap, 'encode_full', False, 'Whether to encode and store the full data set')
    _bool_flag(ap, 'validation', False, 'Whether to use a validation set and print validation loss')
    ap.add_argument('--input_format', default='binary', help='["binary", "tfidf"]')
    ap.add_argument('--label', default='category_publish_name', help='["category_publish_name", "story"]')
    _bool_flag(ap, 'save_tsv', False, 'Whether to save data in tsv format')
    ap.add_argument('--train_row', type=int, default=8000)
    ap.add_argument('--validate_row', type=int, default=2000)
    _bool_flag(ap, 'restore_previous_data', False, 'restore previous data corresponding to model name')
    ap.add_argument('--min_df', type=float, default=0.0)
    ap.add_argument('--max_df', type=float, default=0.99)
    ap.add_argument('--max_features', type=int, default=10000)
    ap.add_argument('--model_name', default='')
    _bool_flag(ap, 'restore_previous_model', False, 'restore previous model corresponding to model name')
    ap.add_argument('--seed', type=int, default=-1)
    ap.add_argument('--compress_factor', type=int, default=20)
    ap.add_argument('--corr_type', default='masking', help='["none", "masking", "salt_and_pepper", "decay"]')
    ap.add_argument('--corr_frac', type=float, default=0.3)
    ap.add_argument('--xavier_init', type=int, default=1)
    ap.add_argument('--enc_act_func', default='sigmoid')
    ap.add_argument('--dec_act_func', default='sigmoid')
    ap.add_argument('--main_dir', default='')
    ap.add_argument('--loss_func', default='cross_entropy')
    ap.add_argument('--opt', default='gradient_descent', help='["gradient_descent", "ada_grad", "momentum", "adam"]')
    ap.add_argument('--learning_rate', type=float, default=0.1)
    ap.add_argument('--momentum', type=float, default=0.5)
    ap.add_argument('--num_epochs', type=int, default=50)
    ap.add_argument('--batch_size', type=float, default=0.1)
    ap.add_argument('--alpha', type=float, default=1.0)
    ap.add_argument('--triplet_strategy', default='batch_all')
    # additive
    ap.add_argument('--data_path', default='datasets/uci_news.snappy.parquet')
    ap.add_argument('--synthetic', type=int, default=0, help='train on N synthetic articles instead of reading --data_path')
    ap.add_argument('--rng_mode', default='device', choices=['numpy', 'device'],
                    help="'device': Philox masking + device permutation (default); 'numpy': the reference's host NumPy RNG stream")
    return ap


def apply_env_overrides(flags):
    """Same-named environment variables (loaded from .env) override the flags (reference main_autoencoder.py:13-17,36-92)."""
    dot_env_path = _script_path / '.env'
    if dot_env_path.exists():
        try:
            import dotenv
            print('.env found, will override all flags using values in .env')
            dotenv.load_dotenv(dot_env_path)
        except ImportError:
            pass
    for k, cast in _ENV_OVERRIDES.items():   # the reference's fixed list (main_autoencoder.py:75-92), with the right keys for corr_*
        if k in os.environ and hasattr(flags, k):
            setattr(flags, k, cast(os.environ[k]))
    return flags


def _env_bool(raw):
    """The reference sets a boolean flag to True when the variable merely exists; here '0' / 'false' / 'no' / '' mean False."""
    return str(raw).strip().lower() not in ('', '0', 'false', 'no', 'off')


_ENV_OVERRIDES = {'model_name': str, 'restore_previous_model': _env_bool, 'seed': int, 'compress_factor': int, 'corr_type': str,
                  'corr_frac': float, 'xavier_init': int, 'enc_act_func': str, 'dec_act_func': str, 'main_dir': str, 'loss_func': str,
                  'opt': str, 'learning_rate': float, 'momentum': float, 'num_epochs': int, 'batch_size': float, 'alpha': float,
                  'triplet_strategy': str}


def check_flags(F):
    assert 0. <= F.min_df <= 1.
    assert 0. <= F.max_df <= 1.
    assert F.max_features >= 1
    assert F.enc_act_func in ['sigmoid', 'tanh']
    assert F.dec_act_func in ['sigmoid', 'tanh', 'none']
    assert F.corr_type in ['masking', 'salt_and_pepper', 'decay', 'none']
    assert 0. <= F.corr_frac <= 1.
    assert F.loss_func in ['cross_entropy', 'mean_squared', 'cosine_proximity']
    assert F.opt in ['gradient_descent', 'ada_grad', 'momentum', 'adam']
    assert F.verbose_step > 0
    assert F.triplet_strategy in ['batch_all', 'batch_hard', 'none']
    assert F.input_format in ['binary', 'tfidf']
    assert F.label in ['category_publish_name', 'story']
    if F.input_format == 'tfidf':
        assert F.loss_func in ['mean_squared', 'cosine_proximity']
    if F.main_dir == '':
        F.main_dir = F.model_name
    return F


_LABELS = ('category_publish_name', 'story')
_TSV_LABEL_COLUMNS = ['label_story', 'label_category_publish_name', 'title', 'story', 'category_publish_name']


def prepare_uci(F, model=None):
    """main_autoencoder.py:177-244 with pandas >= 2 fixes (sort by the article_id column, no DataFrame.append): vectorise the
    newest train_row + validate_row articles and, when a model is given, write the data_dir cache the reference writes
    (same file names and formats) so that --restore_previous_data finds it."""
    import joblib
    import pandas as pd
    from sklearn.feature_extraction.text import CountVectorizer, TfidfTransformer
    from dae_rnn_news_recommendation_b200.io_formats import save_file
    df = pd.read_parquet(F.data_path)
    if 'article_id' in df.columns:
        df = df.set_index('article_id', drop=False)
        df.index.name = None
    df = df.sort_index(ascending=False)
    df['label_story'] = pd.factorize(df.story)[0]
    cat = df.category_publish_name.apply(lambda s: s.lstrip('即時') if isinstance(s, str) else s)
    df['label_category_publish_name'] = pd.factorize(cat)[0]
    if F.triplet_strategy != 'none':
        # rows without the selected label cannot be mined (reference main_autoencoder.py:181-201 keeps label_<label>_valid == 1):
        # pd.factorize gives them -1, which would otherwise act as one large shared class
        df = df.loc[df['label_' + F.label] >= 0]
    n_tr, n_va = F.train_row, F.validate_row
    df = df.iloc[0:n_tr + n_va].sample(frac=1)
    df = df.sort_values('article_id') if 'article_id' in df.columns else df.sort_index()
    def df_bound(v):   # sklearn >= 1.2 validates: a proportion is a float in [0, 1], a document count an int >= 1
        return float(v) if v <= 1 else int(v)
    cv = CountVectorizer(stop_words='english', min_df=df_bound(F.min_df), max_df=df_bound(F.max_df), max_features=F.max_features,
                         binary=False)
    X = cv.fit_transform(df.main_content[0:n_tr])
    Xv = cv.transform(df.main_content[n_tr:n_tr + n_va])
    tf = TfidfTransformer()
    Xt, Xtv = tf.fit_transform(X), tf.transform(Xv)
    d = {'articles': df.iloc[0:n_tr], 'articles_validate': df.iloc[n_tr:n_tr + n_va], 'tfidf': (Xt, Xtv),
         'count_vectorizer': cv, 'tfidf_transformer': tf}
    for lab in _LABELS:
        d['label_' + lab] = (df['label_' + lab][0:n_tr], df['label_' + lab][n_tr:n_tr + n_va])
    if model is not None:
        dd = model.data_dir
        save_file(d['articles'], dd + 'article.snappy.parquet')
        save_file(d['articles_validate'], dd + 'article_validate.snappy.parquet')
        for lab in _LABELS:
            save_file(d['label_' + lab][0], dd + 'article_label_%s.pkl' % lab)
            save_file(d['label_' + lab][1], dd + 'article_label_%s_validate.pkl' % lab)
        save_file(X, dd + 'article_count_vectorized.npz')
        save_file(Xv, dd + 'article_count_vectorized_validate.npz')
    X.data = np.ones(len(X.data), dtype=X.data.dtype)      # binary bag of words (main_autoencoder.py:234-235)
    Xv.data = np.ones(len(Xv.data), dtype=Xv.data.dtype)
    d['binary'] = (X, Xv)
    if model is not None:
        save_file(X, dd + 'article_binary_count_vectorized.npz')
        save_file(Xv, dd + 'article_binary_count_vectorized_validate.npz')
        save_file(Xt, dd + 'article_tfidf_vectorized.npz')
        save_file(Xtv, dd + 'article_tfidf_vectorized_validate.npz')
        joblib.dump(cv, dd + 'count_vectorizer.joblib')
        joblib.dump(tf, dd + 'tfidf_transformer.joblib')
    return d


def restore_uci(model):
    """--restore_previous_data (main_autoencoder.py:161-175): read back the cache of an earlier run with the same model name."""
    import joblib
    from dae_rnn_news_recommendation_b200.io_formats import read_file
    dd = model.data_dir
    d = {'articles': read_file(dd + 'article.snappy.parquet'), 'articles_validate': read_file(dd + 'article_validate.snappy.parquet'),
         'binary': (read_file(dd + 'article_binary_count_vectorized.npz'), read_file(dd + 'article_binary_count_vectorized_validate.npz')),
         'tfidf': (read_file(dd + 'article_tfidf_vectorized.npz'), read_file(dd + 'article_tfidf_vectorized_validate.npz')),
         'count_vectorizer': joblib.load(dd + 'count_vectorizer.joblib'), 'tfidf_transformer': joblib.load(dd + 'tfidf_transformer.joblib')}
    for lab in _LABELS:
        d['label_' + lab] = (read_file(dd + 'article_label_%s.pkl' % lab, data_type='pandas_series'),
                             read_file(dd + 'article_label_%s_validate.pkl' % lab, data_type='pandas_series'))
    return d


def save_tsv(model, d, enc, enc_v):
    """--save_tsv (main_autoencoder.py:292-301): the projector-ready TSV set under tsv_dir."""
    from dae_rnn_news_recommendation_b200.io_formats import save_file
    td = model.tsv_dir
    if d is not None:
        for name in ('tfidf', 'binary'):
            stem = 'article_tfidf_vectorized' if name == 'tfidf' else 'article_binary_count_vectorized'
            save_file(d[name][0], td + stem + '.tsv')
            save_file(d[name][1], td + stem + '_validate.tsv')
        cols = [c for c in _TSV_LABEL_COLUMNS if c in d['articles'].columns]
        save_file(d['articles'][cols], td + 'article_label.tsv')
        save_file(d['articles_validate'][cols], td + 'article_label_validate.tsv')
    save_file(np.asarray(enc), td + 'article_encoded.tsv')
    save_file(np.asarray(enc_v), td + 'article_encoded_validate.tsv')


def prepare_synthetic(F):
    from dae_rnn_news_recommendation_b200.synth import make_sparse, make_labels
    n = F.synthetic
    X = make_sparse(n, F.max_features, 100, 'binary' if F.input_format == 'binary' else 'tfidf', seed=max(F.seed, 0))
    lab = make_labels(n, 4, seed=max(F.seed, 0))
    nv = min(F.validate_row, n // 5)
    return X[:n - nv], X[n - nv:], lab[:n - nv], lab[n - nv:]


def evaluate(F, model, trX, vlX, trL, vlL, enc, enc_v, max_rows=20000):
    """Reference main_autoencoder.py:307-360: cosine similarity of the input space and of the embeddings, the related / unrelated
    comparison for the label in use, and the nearest article of the first rows.  N x N lives on the GPU only; sets larger than
    `max_rows` (the reference runs 8000 / 2000 rows) are skipped."""
    from dae_rnn_news_recommendation_b200 import helpers
    out = {}
    print('calculate similarity')
    in_metric = 'cosine' if F.input_format == 'binary' else 'linear kernel'   # tf-idf rows are already l2-normalised (:314)
    in_name = 'binary_count' if F.input_format == 'binary' else 'tfidf'
    suffix = '(Category)' if F.label == 'category_publish_name' else '(Story)'
    for split, X, E, lab in (('', trX, enc, trL), ('_validate', vlX, enc_v, vlL)):
        if X is None or X.shape[0] < 2 or X.shape[0] > max_rows:
            print('similarity%s skipped: %s rows' % (split, None if X is None else X.shape[0]))
            continue
        for name, data, metric in ((in_name, X, in_metric), ('encoded', E, 'cosine')):
            sim = helpers.pairwise_similarity(data, metric=metric, to_host=False)
            key = 'similarity_boxplot_%s%s%s' % (name, split, suffix)
            out[key] = helpers.visualize_pairwise_similarity(lab, sim, plot='boxplot', title=key, save_path=model.plot_dir + key + '.png')
            print('%s: AUROC %.4f  related median %.4f  unrelated median %.4f' % (
                key, out[key]['auroc'], out[key]['related'].get('median', float('nan')), out[key]['unrelated'].get('median', float('nan'))))
            del sim
        idx, score = helpers.nearest_neighbors(E, metric='cosine')
        out['nearest' + split] = (idx, score)
        for i in range(min(3, len(idx))):
            print('article %d%s: most similar %d (cosine %.4f)' % (i, split, idx[i], score[i]))
    print('calculate similarity done')
    return out


def main(argv=None):
    F = check_flags(apply_env_overrides(build_parser().parse_args(argv)))
    print(__file__ + ': Start')
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder, utils
    model = DenoisingAutoencoder(
        seed=F.seed, model_name=F.model_name, compress_factor=F.compress_factor, enc_act_func=F.enc_act_func,
        dec_act_func=F.dec_act_func, xavier_init=F.xavier_init, corr_type=F.corr_type, corr_frac=F.corr_frac,
        loss_func=F.loss_func, main_dir=F.main_dir, opt=F.opt, learning_rate=F.learning_rate, momentum=F.momentum,
        verbose=F.verbose, verbose_step=F.verbose_step, num_epochs=F.num_epochs, batch_size=F.batch_size, alpha=F.alpha,
        triplet_strategy=F.triplet_strategy, rng_mode=F.rng_mode)
    data = None
    if F.synthetic:
        trX, vlX, trL, vlL = prepare_synthetic(F)
    else:
        data = restore_uci(model) if F.restore_previous_data else prepare_uci(F, model)
        (trX, vlX), (trL, vlL) = data[F.input_format], data['label_' + F.label]
        trX, vlX, trL, vlL = trX.astype(np.float32), vlX.astype(np.float32), np.asarray(trL), np.asarray(vlL)
    print('fit')
    model.fit(train_set=trX, validation_set=vlX if F.validation else None, train_set_label=trL,
              validation_set_label=vlL if F.validation else None, restore_previous_model=F.restore_previous_model)
    with open(model.parameter_file, 'a+') as fh:
        for k in ('train_row', 'validate_row', 'input_format', 'label', 'restore_previous_data', 'restore_previous_model'):
            print('{}={}'.format(k, getattr(F, k)), file=fh)
    print('fit done')
    # inputs are decayed by (1 - corr_frac) at inference (reference main_autoencoder.py:289-290)
    enc = model.transform(utils.decay_noise(trX, F.corr_frac), name='article_encoded', save=F.encode_full)
    enc_v = model.transform(utils.decay_noise(vlX, F.corr_frac), name='article_encoded_validate', save=F.encode_full)
    print('encoded: train %s validate %s (train_time of the last epoch: %.3f s)' % (enc.shape, enc_v.shape, model.train_time or 0.0))
    if F.save_tsv:
        save_tsv(model, data, enc, enc_v)
    model.evaluation = evaluate(F, model, trX, vlX, trL, vlL, enc, enc_v)
    print(__file__ + ': End')
    return model


if __name__ == '__main__':
    main()
