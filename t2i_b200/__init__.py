"""Import alias: the product package lives in ``text-to-image_b200/`` (not a Python identifier);
this package extends its ``__path__`` there so that everything imports as ``t2i_b200.<module>``."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "text-to-image_b200"))
