"""The reference's recipe on its own data (BASELINE.json configs[0]: UCI news, 8000 train / 2000 validate, 10 000 features, H=500,
B=800, masking 0.3, sigmoid/sigmoid, CE, SGD 0.1, 50 epochs) on one B200: articles/s of `fit` (the reference's train_time window)
and the evaluation it ends with -- category / story AUROC of binary-count cosine, tf-idf linear kernel and the embeddings of the
plain DAE (`none`) and the triplet DAE (`batch_all`).  One JSON line.   python tools/uci_quality.py [epochs]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from sklearn.feature_extraction.text import TfidfTransformer
from helpers import load_uci_c1
from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder, utils
from dae_rnn_news_recommendation_b200 import helpers

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 50
os.chdir(os.environ.get('TMPDIR', '/tmp'))
d = load_uci_c1()
tf = TfidfTransformer().fit(d['train_counts'])
inputs = {'train': {'binary_count': (d['train'], 'cosine'), 'tfidf': (tf.transform(d['train_counts']).astype(np.float32), 'linear kernel')},
          'validate': {'binary_count': (d['validate'], 'cosine'), 'tfidf': (tf.transform(d['validate_counts']).astype(np.float32), 'linear kernel')}}
res = {'config': 'C1 UCI news 8000x10000 binary, H=500, B=800, masking 0.3, CE, SGD 0.1, %d epochs, rng_mode=%s' % (epochs, os.environ.get('DAE_RNG_MODE', 'device')), 'fit': {}, 'auroc': {}}


def auroc_all(split, name, data, metric):
    sim = helpers.pairwise_similarity(data, metric=metric, to_host=False)
    for lab in ('category_publish_name', 'story'):
        r = helpers.visualize_pairwise_similarity(d['%s_label_%s' % (split, lab)], sim)
        res['auroc']['%s/%s/%s' % (split, lab, name)] = round(r['auroc'], 5)


for split in ('train', 'validate'):
    for name, (data, metric) in inputs[split].items():
        auroc_all(split, name, data, metric)
for strategy in ('none', 'batch_all'):
    m = DenoisingAutoencoder(model_name='uci_' + strategy, main_dir='uci_' + strategy, compress_factor=20, enc_act_func='sigmoid',
                             dec_act_func='sigmoid', loss_func='cross_entropy', corr_type='masking', corr_frac=0.3, opt='gradient_descent',
                             learning_rate=0.1, num_epochs=epochs, batch_size=0.1, alpha=1, triplet_strategy=strategy, seed=0, verbose=False,
                             rng_mode=os.environ.get('DAE_RNG_MODE', 'device'))   # 'numpy' = the reference's host RNG stream (~6 ms/epoch of np.random)
    t0 = time.time()
    m.fit(d['train'], None, d['train_label_category_publish_name'])
    wall = time.time() - t0
    hist = np.concatenate(m.history)
    res['fit'][strategy] = {'fit_wall_s': round(wall, 3), 'last_epoch_train_time_s': round(float(m.train_time), 5),
                            'articles_per_s_last_epoch': round(8000 / float(m.train_time), 1),
                            'cost_first_last': [float(hist[0, 0]), float(hist[-1, 0])]}
    for split in ('train', 'validate'):
        auroc_all(split, 'encoded_' + strategy, m.transform(utils.decay_noise(d[split], 0.3)), 'cosine')
print(json.dumps(res))
