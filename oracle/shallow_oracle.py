"""CPU restatement of the reference's Transformer_torch/Transformer_EEG.py `ShallowConvNet`: the FIRST
step of that row (oracle + pins); no CUDA path exists for it yet.

TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/): nothing under eav_b200/ imports this file.

forward (Transformer_EEG.py:122-142), eval mode unless `train` (then explicit keep-masks replace the dropouts):
  conv      Conv2d(1, 40, (1, 13), bias=False), no padding            (B,1,30,500) -> (B,40,30,488)      :116
  embedding 40 x Linear(30, 1, bias=False), one per temporal filter   -> v (B, 488, 40)                  :14-34
  12 x TransformerLayer(40, heads=1, qkv=40)                                                              :87-103
      a = softmax(Q K^T / sqrt(40)) V + V, Q/K/V = Linear(40,40,bias=False)                               :36-69
      v = v + drop(LayerNorm(a));  v = v + drop(LayerNorm(Linear(160,40)(drop(relu(Linear(40,160)(v))))))
  BatchNorm2d(40) on (B,40,1,488) -> square -> AvgPool((1,35), stride (1,7)) -> log(clamp(., 1e-7, 1e4))
  dropout -> flatten (2600) -> Linear(2600, nb_classes, bias=False) -> softmax
Parameter names are the reference's state_dict keys.
"""
import torch
import torch.nn.functional as F

N_LAYERS = 12


def shallow_forward(sd, x, train=False, masks=None, bn_momentum=0.1, bn_eps=1e-5):
    """sd: {state_dict key: tensor} (BatchNorm running stats are updated in place in train mode);
    x (B, 1, 30, 500); masks: in train mode a list of keep-masks scaled by 1/(1-p), consumed in call order
    (per layer: after norm1, inside the FFN, after norm2; then the final dropout).  Returns probabilities."""
    it = iter(masks or ())

    def drop(t):
        return t * next(it) if train else t

    h = F.conv2d(x, sd["conv.weight"])                                              # (B, 40, 30, T')
    w_emb = torch.stack([sd[f"embedding.value_proj.{i}.weight"][0] for i in range(40)])   # (40, 30)
    v = torch.einsum("bfct,fc->btf", h, w_emb)                                      # (B, T', 40)
    for l in range(N_LAYERS):
        p = f"transformer.{l}."
        q = v @ sd[p + "attn.W_q.weight"].T
        k = v @ sd[p + "attn.W_k.weight"].T
        val = v @ sd[p + "attn.W_v.weight"].T
        att = F.softmax(q @ k.transpose(-1, -2) / (40 ** 0.5), dim=-1)
        a = att @ val + val
        v = v + drop(F.layer_norm(a, (40,), sd[p + "norm1.weight"], sd[p + "norm1.bias"]))
        f = F.relu(F.linear(v, sd[p + "ffn.net.0.weight"], sd[p + "ffn.net.0.bias"]))
        f = F.linear(drop(f), sd[p + "ffn.net.3.weight"], sd[p + "ffn.net.3.bias"])
        v = v + drop(F.layer_norm(f, (40,), sd[p + "norm2.weight"], sd[p + "norm2.bias"]))
    z = v.permute(0, 2, 1).unsqueeze(2)                                             # (B, 40, 1, T')
    z = F.batch_norm(z, sd["bn.running_mean"], sd["bn.running_var"], sd["bn.weight"], sd["bn.bias"], train,
                     bn_momentum, bn_eps)
    z = F.avg_pool2d(torch.square(z), (1, 35), stride=(1, 7))
    z = torch.log(torch.clamp(z, 1e-7, 1e4)).squeeze(2)
    z = drop(z).flatten(1)
    return F.softmax(F.linear(z, sd["fc.weight"]), dim=1)


def loss_fn(probs, y):
    """nn.CrossEntropyLoss applied to the softmax OUTPUT (Transformer_EEG.py:170,190), like EEGNet_tor."""
    return F.cross_entropy(probs, y)


def renorm_fc_(sd, maxnorm=0.5):
    """After every optimizer step (Transformer_EEG.py:196-199): torch.renorm(fc.weight, p=2, dim=0, maxnorm=0.5)."""
    w = sd["fc.weight"]
    n = w.norm(dim=1, keepdim=True)
    w.mul_(torch.where(n > maxnorm, maxnorm / (n + 1e-7), torch.ones_like(n)))
