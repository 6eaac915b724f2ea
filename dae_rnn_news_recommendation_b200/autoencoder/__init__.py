from .autoencoder import DenoisingAutoencoder  # noqa: F401
from .autoencoder_triplet import DenoisingAutoencoderTriplet  # noqa: F401
