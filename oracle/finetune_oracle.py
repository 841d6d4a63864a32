"""TEST INFRASTRUCTURE ONLY (the checker, never the product).

CPU restatement, in plain fp32 PyTorch, of the CCD recognition / fine-tuning path (BASELINE
config 5): `DINO_Finetune.forward_train` / `forward_test` of TongkunGuan/CCD (reference commit e3fa0b5):

    img [N,3,32,128] -> ViT encoder (final-norm tokens [N,256,E])            Dino/model/dino_vision.py:199-203
                     -> Mlp E -> 512 -> 512 (GELU)                            dino_vision.py:117-133,164
                     -> NRTRDecoder: 6 pre-LN layers (masked self-attention over the T = 25 target tokens,
                        cross-attention over the 256 encoder tokens, FFN 512 -> 256 -> 512), final LN, Linear 512 -> 92
                                                                              Dino/decoder/nrtr_decoder.py:42-152
                     -> TFLoss: cross-entropy of outputs[:, :-1] against targets[:, 1:], PAD ignored, mean
                                                                              Dino/loss/ce_loss.py:94-128

Functional over a flat {name: tensor} state dict with the reference's parameter names; every dropout is the identity
(the parity configuration: `model.eval()` / p = 0, like drop_path_rate = 0 for the pretraining oracle).
Pinned against the UNMODIFIED reference executed in the build container (tests/test_finetune_oracle.py, needs
/root/reference) and by the committed fixtures tests/golden/finetune_vit_tiny_b4.npz / finetune_vit_small_b3.npz
(tests/golden/make_golden_finetune.py).
Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this file.
"""

import numpy as np
import torch
import torch.nn.functional as F

import ccd_oracle as O

D_MODEL, N_HEAD, D_K, D_INNER, N_LAYERS = 512, 8, 64, 256, 6     # Dino/configs/CCD_vision_model_ARD.yaml:64-76
NUM_CLASSES, START_IDX, PAD_IDX, MAX_SEQ_LEN = 93, 91, 92, 25     # AttnConvertor(DICT90, with_unknown): convertor/attn.py:43-66
DEC_LN_EPS = 1e-5                                                 # TFDecoderLayer norms: nn.LayerNorm default eps
FINAL_LN_EPS = 1e-6                                               # NRTRDecoder.layer_norm (nrtr_decoder.py:77)


def position_table(n_position=200, d_hid=D_MODEL):
    """PositionalEncoding._get_sinusoid_encoding_table (transformer_module.py:138-149): float32 arithmetic as there."""
    denominator = torch.Tensor([1.0 / np.power(10000, 2 * (j // 2) / d_hid) for j in range(d_hid)]).view(1, -1)
    tab = torch.arange(n_position).unsqueeze(-1).float() * denominator
    tab[:, 0::2] = torch.sin(tab[:, 0::2])
    tab[:, 1::2] = torch.cos(tab[:, 1::2])
    return tab                                                    # [n_position, d_hid]


def encoder_mlp(sd, x):
    """DINO_Finetune.encoder = Mlp(E, 512, 512) (dino_vision.py:127-133), dropout = identity."""
    h = O.gelu(x @ sd["encoder.fc1.weight"].t() + sd["encoder.fc1.bias"])
    return h @ sd["encoder.fc2.weight"].t() + sd["encoder.fc2.bias"]


def mha(sd, p, q_in, kv_in, mask):
    """MultiHeadAttention.forward (transformer_module.py:72-96): bias-free projections, softmax(q k^T / sqrt(d_k)) with
    masked_fill(mask == 0, -inf) (:26-31), output projection `fc`.  mask broadcastable to [N, H, Tq, Tk] or None."""
    n, tq, _ = q_in.shape
    tk = kv_in.shape[1]
    q = (q_in @ sd[p + "linear_q.weight"].t()).view(n, tq, N_HEAD, D_K).transpose(1, 2)
    k = (kv_in @ sd[p + "linear_k.weight"].t()).view(n, tk, N_HEAD, D_K).transpose(1, 2)
    v = (kv_in @ sd[p + "linear_v.weight"].t()).view(n, tk, N_HEAD, D_K).transpose(1, 2)
    s = torch.matmul(q / D_K ** 0.5, k.transpose(2, 3))
    if mask is not None:
        s = s.masked_fill(mask == 0, float("-inf"))
    o = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(n, tq, N_HEAD * D_K)
    return o @ sd[p + "fc.weight"].t()


def target_mask(trg_seq):
    """get_pad_mask & get_subsequent_mask (nrtr_decoder.py:82-96): key j visible to query i iff j <= i and trg[j] != PAD."""
    t = trg_seq.shape[1]
    pad = (trg_seq != PAD_IDX).unsqueeze(-2)                      # [N,1,T]
    sub = (1 - torch.triu(torch.ones((t, t)), diagonal=1)).unsqueeze(0).bool()
    return (pad & sub).unsqueeze(1)                               # [N,1,T,T]


def decoder_hidden(sd, trg_seq, src):
    """NRTRDecoder._attention (nrtr_decoder.py:98-116) with the default pre-LN operation order
    (transformer_layers.py:147-159); src_mask is None (`_get_mask` with img_metas=None, :118-131)."""
    emb = sd["decoder.trg_word_emb.weight"][trg_seq]              # nn.Embedding(93, 512, padding_idx=92)
    x = emb + position_table()[: trg_seq.shape[1]].unsqueeze(0)
    m = target_mask(trg_seq)
    for l in range(N_LAYERS):
        p = f"decoder.layer_stack.{l}."
        y = F.layer_norm(x, (D_MODEL,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], DEC_LN_EPS)
        x = x + mha(sd, p + "self_attn.", y, y, m)
        y = F.layer_norm(x, (D_MODEL,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], DEC_LN_EPS)
        x = x + mha(sd, p + "enc_attn.", y, src, None)
        y = F.layer_norm(x, (D_MODEL,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], DEC_LN_EPS)
        h = O.gelu(y @ sd[p + "mlp.w_1.weight"].t() + sd[p + "mlp.w_1.bias"])             # transformer_module.py:118-124
        x = x + (h @ sd[p + "mlp.w_2.weight"].t() + sd[p + "mlp.w_2.bias"])
    return F.layer_norm(x, (D_MODEL,), sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"], FINAL_LN_EPS)


def classify(sd, hidden):
    return hidden @ sd["decoder.classifier.weight"].t() + sd["decoder.classifier.bias"]    # [.., 92]


def tf_loss(logits, targets):
    """TFLoss (ce_loss.py:94-128): flatten, outputs[:, :-1] vs targets[:, 1:], ignore_index = PAD, reduction 'mean'."""
    out = logits[:, :-1, :].reshape(-1, logits.shape[-1])
    tgt = targets[:, 1:].reshape(-1)
    return F.cross_entropy(out, tgt, ignore_index=PAD_IDX, reduction="mean")


def finetune_forward_train(sd, arch, img, targets):
    """DINO_Finetune.forward_train (dino_vision.py:205-231): returns (loss, logits [N,T,92], encoder memory [N,256,512])."""
    feat, _ = O.vit_forward(sd, "backbone.", img, arch)
    mem = encoder_mlp(sd, feat)
    logits = classify(sd, decoder_hidden(sd, targets, mem))
    return tf_loss(logits, targets), logits, mem


def finetune_forward_test(sd, arch, img, max_seq_len=MAX_SEQ_LEN):
    """DINO_Finetune.forward_test -> NRTRDecoder.forward_test (nrtr_decoder.py:154-178): greedy decoding that re-runs the
    whole decoder on the growing sequence; returns the per-step softmax [N, max_seq_len, 92]."""
    feat, _ = O.vit_forward(sd, "backbone.", img, arch)
    mem = encoder_mlp(sd, feat)
    n = img.shape[0]
    seq = torch.full((n, max_seq_len + 1), PAD_IDX, dtype=torch.long)
    seq[:, 0] = START_IDX
    outs = []
    for step in range(max_seq_len):
        hid = decoder_hidden(sd, seq, mem)
        prob = torch.softmax(classify(sd, hid[:, step, :]), dim=-1)
        outs.append(prob)
        seq[:, step + 1] = prob.argmax(dim=-1)
    return torch.stack(outs, dim=1)


def synthetic_targets(n, seed=0, max_seq_len=MAX_SEQ_LEN):
    """Random label tensors framed like AttnConvertor.str2tensor (convertor/attn.py:87-105): BOS, 1..max_seq_len-2 characters
    (0..89), EOS (= BOS index 91), PAD = 92 up to max_seq_len."""
    g = torch.Generator().manual_seed(seed)
    t = torch.full((n, max_seq_len), PAD_IDX, dtype=torch.long)
    for i in range(n):
        ln = int(torch.randint(1, max_seq_len - 1, (1,), generator=g))
        t[i, 0] = START_IDX
        t[i, 1:1 + ln] = torch.randint(0, 90, (ln,), generator=g)
        t[i, 1 + ln] = START_IDX                                  # start_end_same=True: EOS shares the BOS index
    return t
