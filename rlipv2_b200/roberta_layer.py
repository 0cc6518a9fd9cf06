"""RobertaLayer - the post-LN BERT block interleaved with ALIF in the ParSeDA encoder.

Mirror of /root/reference/models/modeling_roberta.py:340-408 (+ RobertaSelfAttention :117-241,
RobertaSelfOutput :245-256, RobertaIntermediate :309-321, RobertaOutput :325-336) with the same
sub-module names, so that `transformer.encoder.roberta_layers.N.*` checkpoints load:

  ext = (1 - mask) * -10000                      (transformers 4.5.1 get_extended_attention_mask)
  a   = softmax(Q K^T / sqrt(64) + ext) V        12 heads x 64
  h1  = LN(dense(a) + x)                         eps = config.layer_norm_eps (1e-5)
  out = LN(dense(gelu(dense(h1))) + h1)

Dropouts (p = 0.1 from the HF config) are active in training mode, as in the reference.
"""
import math

from torch import nn

from . import dense


class RobertaSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def forward(self, hidden_states, extended_mask):
        """extended_mask: [b, 1, 1, t] additive (0 / -10000) or None"""
        h, d = self.num_attention_heads, self.attention_head_size
        q = dense.linear(hidden_states, self.query.weight, self.query.bias)
        k = dense.linear(hidden_states, self.key.weight, self.key.bias)
        v = dense.linear(hidden_states, self.value.weight, self.value.bias)
        key_bias = extended_mask.reshape(extended_mask.shape[0], -1) if extended_mask is not None else None
        # softmax(q k^T / sqrt(d) + mask) -> dropout -> . v  (modeling_roberta.py:185-241): dense.attention
        return dense.attention(q, k, v, h, 1.0 / math.sqrt(d), key_bias, self.dropout.p, self.training,
                               salt=getattr(self, "_rlipv2_salt", 0))


class RobertaSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        h = self.dropout(dense.linear(hidden_states, self.dense.weight, self.dense.bias))
        return dense.add_layer_norm(h, input_tensor, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps)


class RobertaAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = RobertaSelfAttention(config)
        self.output = RobertaSelfOutput(config)

    def forward(self, hidden_states, extended_mask):
        return self.output(self.self(hidden_states, extended_mask), hidden_states)


class RobertaIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        if config.hidden_act != "gelu":
            raise NotImplementedError("roberta-base uses gelu")

    def forward(self, hidden_states):
        return dense.linear_gelu(hidden_states, self.dense.weight, self.dense.bias)


class RobertaOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        h = self.dropout(dense.linear(hidden_states, self.dense.weight, self.dense.bias))
        return dense.add_layer_norm(h, input_tensor, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps)


class RobertaLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.attention = RobertaAttention(config)
        self.intermediate = RobertaIntermediate(config)
        self.output = RobertaOutput(config)

    def forward(self, hidden_states, attention_mask=None):
        """attention_mask: [b, t], 1/True = keep, 0/False = masked (modeling_roberta.py:366-376)."""
        ext = None
        if attention_mask is not None:
            ext = (1.0 - attention_mask[:, None, None, :].to(hidden_states.dtype)) * -10000.0
        attn_out = self.attention(hidden_states, ext)
        return self.output(self.intermediate(attn_out), attn_out)
