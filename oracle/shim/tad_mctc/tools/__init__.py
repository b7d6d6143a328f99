from . import memory
