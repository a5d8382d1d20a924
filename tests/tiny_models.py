# -*- coding: utf-8 -*-
"""seeded tiny random-init HF models of the four BASELINE families (no checkpoints exist offline)"""
import torch


def tiny_config(family, vocab=64, **over):
    from transformers import GPT2Config, LlamaConfig, MistralConfig, MixtralConfig
    if family == 'gpt2':
        cfg = GPT2Config(vocab_size=vocab, n_positions=512, n_embd=64, n_layer=2, n_head=4, bos_token_id=1,
                         eos_token_id=2)
    elif family == 'llama':
        cfg = LlamaConfig(vocab_size=vocab, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                          num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=1024,
                          rms_norm_eps=1e-6, bos_token_id=1, eos_token_id=2, pad_token_id=0)
    elif family == 'mistral':
        cfg = MistralConfig(vocab_size=vocab, hidden_size=512, intermediate_size=512, num_hidden_layers=2,
                            num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=1024,
                            rms_norm_eps=1e-6, sliding_window=None, bos_token_id=1, eos_token_id=2, pad_token_id=0)
    elif family == 'mixtral':
        cfg = MixtralConfig(vocab_size=vocab, hidden_size=256, intermediate_size=256, num_hidden_layers=2,
                            num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=1024,
                            num_local_experts=4, num_experts_per_tok=2, rms_norm_eps=1e-6, sliding_window=None,
                            bos_token_id=1, eos_token_id=2, pad_token_id=0)
    else:
        raise ValueError(family)
    for k, v in over.items():
        setattr(cfg, k, v)
    cfg._attn_implementation = 'eager'
    return cfg


def tiny_hf_model(family, seed=0, dtype=torch.float32, device='cpu', vocab=64, **over):
    from transformers import AutoModelForCausalLM
    torch.manual_seed(seed)
    cfg = tiny_config(family, vocab=vocab, **over)
    model = AutoModelForCausalLM.from_config(cfg, attn_implementation='eager')
    # a larger init makes the random model less degenerate than std=0.02 (sharper logits, clearer argmax margins)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.08)
    return model.to(device=device, dtype=dtype).eval()


def prompts(seed, n, length, vocab):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(3, vocab, (1, length), generator=g) for _ in range(n)]
