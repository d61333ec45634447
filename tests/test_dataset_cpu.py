"""Input pipeline: blim_b200.dataset against the reference's dataloader classes.

Golden: tests/golden/dataset_tiny.npz, written by oracle/make_dataset_golden.py from the UNMODIFIED reference on the
miniature data tree of oracle/dataset_fixture.py with the stub tokenizer.  Integer outputs must be bit-exact.
"""
import os
import types

import numpy as np
import pytest
import torch

from blim_b200 import dataset as D
from oracle import dataset_fixture
from oracle.stub_tokenizer import StubTokenizer

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_tiny.npz"))
FIELDS = ("vtg_ids", "vtg_labels", "vtg_masks", "tvg_ids", "tvg_labels", "tvg_masks")


@pytest.fixture(scope="module")
def data_root(tmp_path_factory):
    root = tmp_path_factory.mktemp("blim_data") / "data"
    dataset_fixture.write(str(root))
    return str(root)


@pytest.mark.parametrize("bos", [None, 7])
@pytest.mark.parametrize("split", ["test", "train"])
@pytest.mark.parametrize("name", sorted(dataset_fixture.FILES))
def test_dataset_matches_reference_golden(data_root, name, split, bos):
    args = types.SimpleNamespace(dataset=name, batch_size_eval=3)
    ds = D.DATASETS[name](args=args, tokenizer=StubTokenizer(bos=bos), split=split, root=data_root)
    key = f"{name}/{split}/bos{bos}"
    assert len(ds) == int(GOLD[f"{key}/n"]) and len(ds) == (6 if split == "test" else 5)
    assert ds.tvg_prefix_length == int(GOLD[f"{key}/tvg_prefix_length"])
    assert list(ds.vids) == list(GOLD[f"{key}/vids"])
    assert [d["text"] for d in ds.data] == list(GOLD[f"{key}/texts"])
    np.testing.assert_array_equal(ds.video_vocab[:, :, ::16].numpy(), GOLD[f"{key}/video_vocab_sub"])
    np.testing.assert_array_equal(ds.video_vocab.double().sum(-1).numpy(), GOLD[f"{key}/video_vocab_sum"])
    items = [ds[i] for i in range(len(ds))]
    for f in FIELDS:
        np.testing.assert_array_equal(np.concatenate([it[f].numpy() for it in items]), GOLD[f"{key}/{f}"])
        np.testing.assert_array_equal(np.array([len(it[f]) for it in items]), GOLD[f"{key}/{f}_len"])
    np.testing.assert_array_equal(np.array([it["tvg_video_labels"] for it in items]), GOLD[f"{key}/tvg_video_labels"])
    batch = ds.collate_fn(items[:3])
    for f in FIELDS:
        v = batch[f]
        assert torch.is_tensor(v) == (split == "train")     # left-padded tensors only for the train split
        got = v.numpy() if torch.is_tensor(v) else np.concatenate([x.numpy() for x in v])
        np.testing.assert_array_equal(got, GOLD[f"{key}/collate_{f}"])
    np.testing.assert_array_equal(batch["tvg_video_labels"].numpy(), GOLD[f"{key}/collate_tvg_video_labels"])


def test_prompt_shape_feeds_the_engine_contract(data_root):
    """One -200 sentinel per prompt; VTG labels cover caption + <|im_end|> + newline, TVG labels cover the image slot +
    <|im_end|> + newline; the missing feature file yields zeros."""
    args = types.SimpleNamespace(dataset="MSRVTT", batch_size_eval=4)
    loader = D.load_data(args, tokenizer=StubTokenizer(), split="test", root=data_root)
    ds = loader.dataset
    n = 0
    for batch in loader:
        for i in range(len(batch["vid"])):
            vi, vl = batch["vtg_ids"][i], batch["vtg_labels"][i]
            ti, tl = batch["tvg_ids"][i], batch["tvg_labels"][i]
            assert int((vi == D.IMAGE_TOKEN_INDEX).sum()) == 1 and int((ti == D.IMAGE_TOKEN_INDEX).sum()) == 1
            assert vl[-2:].tolist() == [151645, 198] and tl[-3:].tolist() == [D.IMAGE_TOKEN_INDEX, 151645, 198]
            assert int((tl != D.IGNORE_INDEX).sum()) == 3
            p = int((ti == D.IMAGE_TOKEN_INDEX).nonzero()[0])
            assert p > ds.tvg_prefix_length
            if batch["vid"][i] == "video4":
                assert float(batch["video"][i].abs().sum()) == 0.0
            n += 1
    assert n == len(ds)


def test_reference_dataloader_live(data_root):
    """When the reference tree is present (build container) regenerate the golden outputs and compare them with the file."""
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    from oracle.make_dataset_golden import reference_outputs
    live = reference_outputs(os.path.dirname(data_root))
    assert sorted(live) == sorted(GOLD.files)
    for k in live:
        np.testing.assert_array_equal(live[k], GOLD[k])
