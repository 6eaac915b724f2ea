"""On-disk formats of the reference's data_dir / tsv_dir caches (reference helpers.py:138-264 and its
call sites main_autoencoder.py:161-244, 292-301).  Host-side I/O only -- no GPU work, no torch import.

    save_file(data, path, format=None, **kw)        ndarray -> csv / tsv / npy;  scipy sparse -> npz (csv / tsv densify first);
                                                    DataFrame -> csv / tsv / parquet / pkl;  Series -> csv / tsv / pkl
    read_file(path, data_type=None, format=None)    data_type in {'numpy', 'scipy', 'pandas_df', 'pandas_series'}; inferred from the
                                                    extension when omitted (npy -> numpy, npz -> scipy, anything else -> pandas_df)

Same names, argument meaning, file layouts and assertion behaviour as the reference, written against the pandas / pyarrow of this
image (the reference's `to_parquet(fname=)`, `Series.to_csv(path=)` and `read_csv(squeeze=)` spellings no longer exist): Series are
stored header-less with their index in column 0, DataFrames with a header row and the index in column 0.
"""
import os
import warnings

import numpy as np
import pandas as pd
import scipy.sparse as sparse

_WRITABLE = {'numpy': ('csv', 'tsv', 'npy'), 'scipy': ('npz',), 'pandas_df': ('csv', 'tsv', 'parquet', 'pkl'),
             'pandas_series': ('csv', 'tsv', 'pkl')}
_READABLE = {'numpy': ('csv', 'tsv', 'npy'), 'scipy': ('csv', 'tsv', 'npz'), 'pandas_df': ('csv', 'tsv', 'parquet', 'pkl'),
             'pandas_series': ('csv', 'tsv', 'pkl')}
_SEP = {'csv': ',', 'tsv': '\t'}


def _extension(path):
    return str(path).lower().split('.')[-1]


def _kind_of(data):
    if isinstance(data, np.ndarray):
        return 'numpy'
    if sparse.issparse(data):
        return 'scipy'
    if isinstance(data, pd.DataFrame):
        return 'pandas_df'
    if isinstance(data, pd.Series):
        return 'pandas_series'
    return None


def save_file(data, path, format=None, **savekwargs):
    path = str(path)
    fmt = _extension(path) if format is None else format
    if sparse.issparse(data) and fmt in _SEP:   # text formats hold the dense matrix (helpers.py:146-147)
        data = data.toarray()
    kind = _kind_of(data)
    assert kind is not None, 'unsupported data type {}'.format(type(data))
    assert fmt in _WRITABLE[kind], 'Shoule be one of following format {}'.format(list(_WRITABLE[kind]))
    if kind == 'numpy':
        if fmt == 'npy':
            np.save(path, data, **savekwargs)
        else:
            np.savetxt(path, data, delimiter=_SEP[fmt], **savekwargs)
    elif kind == 'scipy':
        sparse.save_npz(path, data, **savekwargs)
    elif fmt in _SEP:
        data.to_csv(path, sep=_SEP[fmt], **({'header': False} if kind == 'pandas_series' else {}), **savekwargs)
    elif fmt == 'parquet':
        data.to_parquet(path, **savekwargs)
    else:
        data.to_pickle(path, **savekwargs)


def read_file(path, data_type=None, format=None, **readkwargs):
    path = str(path)
    assert os.path.isfile(path), '[Error] {} is not a file'.format(path)
    fmt = _extension(path) if format is None else format
    if data_type is None:
        data_type = {'npy': 'numpy', 'npz': 'scipy'}.get(fmt, 'pandas_df')
    assert data_type in _READABLE
    assert fmt in _READABLE[data_type]
    if data_type in ('numpy', 'scipy'):
        if fmt == 'npy':
            return np.load(path, **readkwargs)
        if fmt == 'npz':
            return sparse.load_npz(path, **readkwargs)
        dense = np.loadtxt(path, delimiter=_SEP[fmt], **readkwargs)
        return dense if data_type == 'numpy' else sparse.csr_matrix(dense)
    if fmt == 'parquet':
        return pd.read_parquet(path, **readkwargs)
    if fmt == 'pkl':
        return pd.read_pickle(path, **readkwargs)
    with warnings.catch_warnings():   # parse_dates=True on a non-date index only warns; the index is kept as read
        warnings.simplefilter('ignore', UserWarning)
        if data_type == 'pandas_df':
            return pd.read_csv(path, sep=_SEP[fmt], index_col=0, parse_dates=True, **readkwargs)
        frame = pd.read_csv(path, sep=_SEP[fmt], index_col=0, parse_dates=True, header=None, **readkwargs)
    series = frame.iloc[:, 0]            # what `squeeze=True` used to return
    series.name, series.index.name = None, None
    return series
