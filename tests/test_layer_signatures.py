"""CPU: the wrappers keep the reference's constructor / forward signatures (parsed from the reference sources when
/root/reference is mounted; the expected argument lists are also pinned here so the GPU box checks them too)."""
import ast
import inspect
import os

import pytest

REF = "/root/reference/model"
CTOR = ["originalLayer", "kv_cache", "p8_nums", "p6_nums", "reorder_index", "layer_idx"]
FWD = ["hidden_states", "attention_mask", "position_ids", "past_key_value", "output_attentions", "use_cache",
       "cache_position", "position_embeddings"]


def _ours():
    from micromix_b200.qLlamaLayer import QLlamaDecoderLayer
    from micromix_b200.qMixtralLayer import QMixtralDecoderLayer, QMixtralSparseMoeBlock
    from micromix_b200.qQwenLayer import QQwen2DecoderLayer
    return QLlamaDecoderLayer, QQwen2DecoderLayer, QMixtralDecoderLayer, QMixtralSparseMoeBlock


def _args(fn):
    return [p for p in inspect.signature(fn).parameters if p not in ("self", "kwargs")]


def test_pinned_signatures():
    llama, qwen, mixtral, moe = _ours()
    for cls in (llama, qwen, mixtral):
        assert _args(cls.__init__)[:6] == CTOR
    for cls in (llama, qwen):
        assert _args(cls.forward)[:8] == FWD
    assert _args(mixtral.forward)[:9] == FWD[:5] + ["output_router_logits"] + FWD[5:]  # qMixtralLayer.py:119-129
    assert _args(moe.__init__)[:5] == ["originalSparseMoeBlock", "p8_nums", "p6_nums", "reorder_index", "i"]
    from micromix_b200.qLinearLayer import QLinearLayer
    assert _args(QLinearLayer.__init__) == ["originalLayer", "p8_num", "p6_num", "reorder_index", "out_reorder_index"]


def _ref_args(path, cls, fn):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == fn:
                    return [a.arg for a in f.args.args if a.arg != "self"]
    raise AssertionError(f"{cls}.{fn} not found in {path}")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_signatures_match_reference_sources():
    llama, qwen, mixtral, moe = _ours()
    for path, cls, ours in (("qLlamaLayer.py", "QLlamaDecoderLayer", llama), ("qQwenLayer.py", "QQwen2DecoderLayer", qwen),
                            ("qMixtralLayer.py", "QMixtralDecoderLayer", mixtral)):
        ref_ctor = _ref_args(os.path.join(REF, path), cls, "__init__")
        ref_fwd = _ref_args(os.path.join(REF, path), cls, "forward")
        assert _args(ours.__init__)[:len(ref_ctor)] == ref_ctor, (cls, ref_ctor)
        ours_fwd = _args(ours.forward)
        assert [a for a in ref_fwd if a in ours_fwd] == ref_fwd, (cls, ref_fwd, ours_fwd)
    assert _args(moe.__init__)[:5] == _ref_args(os.path.join(REF, "qMixtralLayer.py"), "QMixtralSparseMoeBlock", "__init__")


def test_moe_token_grouping_matches_the_reference_loop():
    """group_tokens_by_expert == the reference's per-expert torch.where (qMixtralLayer.py:437-450), on CPU."""
    import torch
    from micromix_b200.qMixtralLayer import group_tokens_by_expert
    g = torch.Generator().manual_seed(3)
    for tokens, E, k in ((1, 8, 2), (37, 8, 2), (500, 4, 2), (64, 8, 1), (9, 3, 3)):
        scores = torch.rand(tokens, E, generator=g)
        w, sel = torch.topk(scores, k, dim=-1)
        order, tok_sorted, counts = group_tokens_by_expert(sel, E)
        assert sum(counts) == tokens * k and len(counts) == E
        off = 0
        for e in range(E):
            tok, slot = torch.where(sel == e)
            n = counts[e]
            assert n == tok.numel()
            assert torch.equal(tok_sorted[off:off + n], tok)
            assert torch.equal(w.reshape(-1)[order][off:off + n], w[tok, slot])
            off += n


def test_packed_checkpoint_validation_cpu():
    """from_packed() validates format / shapes before anything touches the GPU."""
    import pytest
    import torch
    from micromix_b200.qLinearLayer import PACKED_FORMAT, QLinearLayer
    with pytest.raises(ValueError):
        QLinearLayer.from_packed({"format": "something else"})
    st = {"format": PACKED_FORMAT, "in_features": 512, "out_features": 256, "p4_num": 256, "p6_num": 128, "p8_num": 100,
          "reorder_index": torch.arange(512, dtype=torch.int16), "bias": None}
    with pytest.raises(ValueError):
        QLinearLayer.from_packed(st)
