from .naf import NAF, ImageEncoder, KeyEncoder, QueryEncoder

__all__ = ["NAF", "ImageEncoder", "KeyEncoder", "QueryEncoder"]
