from .cross_view_attention import CrossViewAttention
from .decoder import Decoder
from .encoder import Encoder
from .merger import Merger
from .refiner import Refiner
from .swin_transformer import SwinTransformer

__all__ = ["Encoder", "SwinTransformer", "CrossViewAttention", "Decoder", "Merger", "Refiner"]
