# -*- coding: utf-8 -*-
"""API types of the lookahead generation path, mirroring
/root/reference/lookahead/lookahead/common/lookahead_generation_utils.py (:19-29 config, :32-47 modes, :50-77 output)."""
from dataclasses import dataclass
from enum import Enum
from typing import Dict, Optional, Tuple

import torch


class GenerationMode(str, Enum):
    """generation modes this path can take (reference :32-47 lists HF's full set; only the two reachable from the
    drop-in surface exist here)"""
    GREEDY_SEARCH = 'greedy_search'
    LOOKAHEAD_GENERATION = 'lookahead_generation'
    SAMPLE = 'sample'


@dataclass
class LookaheadDecoderOnlyOutput(object):
    """reference :50-77. `kwargs` carries the per-call lists dls / edls / fts / qts (README :217-233)."""
    sequences: torch.LongTensor = None
    scores: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple] = None
    hidden_states: Optional[Tuple] = None
    kwargs: Optional[Dict] = None

    def __getitem__(self, k):
        return getattr(self, k) if isinstance(k, str) else (self.sequences, self.scores, self.attentions,
                                                            self.hidden_states, self.kwargs)[k]


class LookaheadGenerationConfig(object):
    """defaults of the reference's LookaheadGenerationConfig (:19-29)"""

    def __init__(self, **kwargs):
        self.use_lookahead = kwargs.pop('use_lookahead', False)
        self.debug_lookahead = kwargs.pop('debug_lookahead', False)
        self.decoding_length = kwargs.pop('decoding_length', 63)
        self.branch_length = kwargs.pop('branch_length', 12)
        self.decoding_mode = kwargs.pop('decoding_mode', 'hier')
        self.decoding_kwargs = kwargs.pop('decoding_kwargs', {})
        self.inputs_embeds_position = kwargs.pop('inputs_embeds_position', False)
        for k, v in kwargs.items():
            setattr(self, k, v)
