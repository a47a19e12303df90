# same public surface as the reference's arch/__init__.py:1-3
from .generators import define_Gen
from .discriminators import define_Dis
from .ops import set_grad
