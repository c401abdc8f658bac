from .multimod_decoder import MultiModDecoder
from .decoders import ClassDecoder, MLPDecoder, LogisticDecoder

__all__ = ["MultiModDecoder", "ClassDecoder", "MLPDecoder", "LogisticDecoder"]
