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
