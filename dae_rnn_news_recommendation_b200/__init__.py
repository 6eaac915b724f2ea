"""dae_rnn_news_recommendation_b200: the DAE-with-triplet-loss training / encoding hot path of
louislung/DAE_RNN_News_Recommendation on B200 (sm_100a), behind the reference's own Python API.

    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder

Importing the package does not need a GPU; constructing an engine (fit / transform) does, and fails loudly
if libdae_sm100.so is missing (there is no CPU fallback).
"""
__version__ = '0.1.0'
