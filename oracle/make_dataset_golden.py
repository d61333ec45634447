"""TEST INFRASTRUCTURE -- tests/golden/dataset_tiny.npz: outputs of the UNMODIFIED reference dataloader classes
(/root/reference/dataloader) on the miniature data tree of oracle/dataset_fixture.py with oracle/stub_tokenizer.py.
Run in the build container only:   python oracle/make_dataset_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dataset_fixture, ref_harness  # noqa: E402
from oracle.stub_tokenizer import StubTokenizer  # noqa: E402


def reference_outputs(root_dir):
    ref_harness._install_stubs()
    if ref_harness.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REF_ROOT)
    import dataloader as ref_dl
    out = {}
    cwd = os.getcwd()
    os.chdir(root_dir)   # the reference hard-codes ./data (base_dataset.py:16)
    try:
        for bos in (None, 7):
            tok = StubTokenizer(bos=bos)
            for name in dataset_fixture.FILES:
                for split in ("test", "train"):
                    args = types.SimpleNamespace(dataset=name)
                    ds = getattr(ref_dl, name)(args=args, tokenizer=tok, image_processor=None, split=split)
                    key = f"{name}/{split}/bos{bos}"
                    out[f"{key}/n"] = np.array(len(ds))
                    out[f"{key}/tvg_prefix_length"] = np.array(ds.tvg_prefix_length)
                    out[f"{key}/vids"] = np.array(ds.vids)
                    out[f"{key}/video_vocab_sub"] = ds.video_vocab[:, :, ::16].numpy()          # every 16th column, exact
                    out[f"{key}/video_vocab_sum"] = ds.video_vocab.double().sum(-1).numpy()     # + per-(video, clip) checksums
                    items = [ds[i] for i in range(len(ds))]
                    for f in ("vtg_ids", "vtg_labels", "vtg_masks", "tvg_ids", "tvg_labels", "tvg_masks"):
                        out[f"{key}/{f}"] = np.concatenate([it[f].numpy() for it in items])
                        out[f"{key}/{f}_len"] = np.array([len(it[f]) for it in items])
                    out[f"{key}/tvg_video_labels"] = np.array([it["tvg_video_labels"] for it in items])
                    out[f"{key}/texts"] = np.array([d["text"] for d in ds.data])
                    batch = ds.collate_fn(items[:3])
                    for f in ("vtg_ids", "vtg_labels", "vtg_masks", "tvg_ids", "tvg_labels", "tvg_masks"):
                        v = batch[f]
                        out[f"{key}/collate_{f}"] = v.numpy() if torch.is_tensor(v) else np.concatenate([x.numpy() for x in v])
                    out[f"{key}/collate_tvg_video_labels"] = batch["tvg_video_labels"].numpy()
    finally:
        os.chdir(cwd)
    return out


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        dataset_fixture.write(os.path.join(tmp, "data"))
        res = reference_outputs(tmp)
    path = os.path.join(ROOT, "tests", "golden", "dataset_tiny.npz")
    np.savez_compressed(path, **res)
    print(path, os.path.getsize(path), "bytes,", len(res), "arrays")
