"""TEST INFRASTRUCTURE -- a miniature ./data tree (annotations + feature files) in the four datasets' own formats
(reference dataloader/{msrvtt,didemo,activitynet,lsmdc}.py), shared by the golden generator and the parity test."""
import json
import os

import torch

CAPTIONS = ["a man is talking about a car", "two dogs run across the field.", "someone slices an onion, then fries it",
            "the crowd cheers as the band starts to play", "a girl opens the door and looks outside", "kids are playing football in the rain"]

FILES = {  # dataset -> {split: annotation file}
    "MSRVTT": {"train": "msrvtt_ret_train.json", "test": "msrvtt_ret_test.json"},
    "DiDeMo": {"train": "didemo_ret_train.json", "test": "didemo_ret_test.json"},
    "ActivityNet": {"train": "anet_ret_train.json", "test": "anet_ret_val_1.json"},
    "LSMDC": {"train": "lsmdc_ret_train.json", "test": "lsmdc_ret_test_1000.json"},
}


def annotations(dataset):
    out = []
    for i, cap in enumerate(CAPTIONS):
        # ids deliberately not in sorted order, and one video that owns two captions (vocabulary < items)
        vid = f"video{(7 * i) % 10}" if i != 4 else "video0"
        if dataset == "MSRVTT":
            out.append({"video": f"{vid}.mp4", "caption": f"  {cap} "})
        elif dataset == "DiDeMo":
            out.append({"video": f"{vid}.avi", "caption": [cap, CAPTIONS[(i + 1) % len(CAPTIONS)]]})
        elif dataset == "ActivityNet":
            out.append({"video": f"{vid}.mp4", "caption": [cap + ". ", CAPTIONS[(i + 2) % len(CAPTIONS)]]})
        else:
            out.append({"video": f"movie{i % 2}/{vid}.avi", "caption": cap + " "})
    return out


def write(root, n_clips=4):
    """Creates {root}/{dataset}/... for all four datasets; the feature of `video4` is missing on purpose (zeros at test
    time, dropped from the train split; base_dataset.py:26-31, msrvtt.py:12)."""
    for dataset, files in FILES.items():
        os.makedirs(f"{root}/{dataset}/features", exist_ok=True)
        ann = annotations(dataset)
        for split, name in files.items():
            json.dump(ann, open(f"{root}/{dataset}/{name}", "w"))
        vids = sorted({(a["video"][:-4].split("/")[1] if dataset == "LSMDC" else a["video"].split(".")[0]) for a in ann})
        g = torch.Generator().manual_seed(11)
        for v in vids:
            feat = (torch.randn(n_clips, 64, 1024, generator=g) * 0.5).half()
            if v != "video4":
                torch.save(feat, f"{root}/{dataset}/features/{v}.pth")
