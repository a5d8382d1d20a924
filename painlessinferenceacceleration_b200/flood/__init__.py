# -*- coding: utf-8 -*-
"""FLOOD integration point: the `Spec` speculative-decoding interface of
/root/reference/flood/flood/utils/speculative.py with its lookahead draft on libpia_b200's CUDA kernels."""
from .speculative import Lookahead, Spec  # noqa: F401
