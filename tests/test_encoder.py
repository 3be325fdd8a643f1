"""Transformer encoder presets (fast_forward/encoder/transformer.py) against the unmodified
reference classes (src/fast_forward/encoder/transformer.py) on a tiny randomly initialised BERT
written to a temporary directory — the reference's own tests (tests/test_encoder.py:30-91) need
downloaded checkpoints.  CPU only."""

import importlib.util
import os

import numpy as np
import pytest

REF = os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "fast_forward", "encoder", "transformer.py")
TEXTS = ["hello world", "a much longer piece of text about re ranking with dense vectors", "q"]


@pytest.fixture(scope="module")
def tiny_bert(tmp_path_factory):
    torch = pytest.importorskip("torch")
    transformers = pytest.importorskip("transformers")
    path = tmp_path_factory.mktemp("tiny-bert")
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "[", "]", "q", "d", "a", "hello", "world", "much",
             "longer", "piece", "of", "text", "about", "re", "ranking", "with", "dense", "vectors", "##s"]
    (path / "vocab.txt").write_text("\n".join(words) + "\n")
    torch.manual_seed(0)
    config = transformers.BertConfig(vocab_size=len(words), hidden_size=32, num_hidden_layers=2,
                                     num_attention_heads=2, intermediate_size=64, max_position_embeddings=64)
    transformers.BertModel(config).save_pretrained(path)
    transformers.BertTokenizer(str(path / "vocab.txt")).save_pretrained(path)
    return path


def test_lambda_and_table_encoders():
    from fast_forward.encoder import LambdaEncoder, TableEncoder

    enc = LambdaEncoder(lambda text: np.full(4, len(text), np.float32))
    assert enc(["ab", "abcd"]).tolist() == [[2.0] * 4, [4.0] * 4]
    table = TableEncoder({"x": np.arange(3.0), "y": np.ones(3)})
    assert table(["y", "x"]).tolist() == [[1, 1, 1], [0, 1, 2]]


@pytest.mark.parametrize("preset", ["TransformerEncoder", "TCTColBERTQueryEncoder", "TCTColBERTDocumentEncoder",
                                    "TASBEncoder", "ContrieverEncoder", "BGEEncoder"])
def test_presets_equal_the_reference_classes(tiny_bert, preset):
    if not os.path.exists(REF):
        pytest.skip("reference package not installed")
    import fast_forward.encoder as mine

    spec = importlib.util.spec_from_file_location("_reference_transformer", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ours, theirs = getattr(mine, preset)(tiny_bert), getattr(ref, preset)(tiny_bert)
    got, want = ours(TEXTS), theirs(TEXTS)
    assert got.shape == (3, 32) and got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)
    if preset == "BGEEncoder":
        np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, rtol=1e-5)


def test_pooling_rules(tiny_bert):
    """Without the reference at hand: the pooling of each preset from the model's own states."""
    import torch
    from fast_forward.encoder import ContrieverEncoder, TCTColBERTDocumentEncoder, TransformerEncoder

    base = TransformerEncoder(tiny_bert)
    inputs = base._tokenizer(TEXTS, return_tensors="pt", padding=True, truncation=True)
    with torch.no_grad():
        hidden = base._model(**inputs).last_hidden_state
    np.testing.assert_array_equal(base(TEXTS), hidden[:, 0].numpy())
    mask = inputs["attention_mask"].unsqueeze(-1).float()
    np.testing.assert_allclose(ContrieverEncoder(tiny_bert)(TEXTS), ((hidden * mask).sum(1) / mask.sum(1)).numpy(),
                               rtol=1e-5, atol=1e-6)
    doc = TCTColBERTDocumentEncoder(tiny_bert, max_length=32)
    marked = doc._tokenizer(doc._get_tokenizer_inputs(TEXTS), return_tensors="pt", **doc._tokenizer_call_args)
    assert marked["input_ids"][0, :4].tolist() == [2, 5, 8, 6]  # [CLS] [ d ]: the 4 positions the mean skips
