from .count import exp_count
from .d3 import cn_d3
