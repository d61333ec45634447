"""Input pipeline of the scoring path: the reference's dataloader package restated on top of the
engine, same names and same outputs.

    RetrievalDataset            dataloader/base_dataset.py:11-163   (BaseDataset)
    MSRVTT / DiDeMo / ActivityNet / LSMDC   dataloader/{msrvtt,didemo,activitynet,lsmdc}.py  (annotation parsing)
    load_data(args, tokenizer, split)       dataloader/__init__.py:8-19 (evaluation loader; no DistributedSampler:
                                            evaluation() reads the whole set on every rank, retrieval_utils.py:182-193)
    stage_corpus(model, dataset)            feature files -> one pinned host tensor -> device corpus, video_vocab
                                            (feature.mean(1), base_dataset.py:33-37) built by the engine's kernel

What is produced per item (base_dataset.py:60-114): prompt token ids with one -200 image sentinel, labels = ids with the
prompt part set to -100, masks = ids != pad.  The chat template is the reference's "qwen_2" ChatML conversation
(conversation.py:90-100,440-449) written out as a string builder; the tokenizer is whatever object the caller passes
(callable returning .input_ids, with .pad_token_id / .bos_token_id), exactly like the reference.
"""
import copy
import glob
import json

import torch

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"

_SYSTEM = "<|im_start|>system\nYou are a helpful assistant."
_ROLES = ("<|im_start|>user", "<|im_start|>assistant")
_SEP = "<|im_end|>"

VTG_PROMPTS = {"DiDeMo": "Describe this video in detail.", "ActivityNet": "Describe this video in detail.",
               "LSMDC": "Describe this video in one sentence.", "MSRVTT": "Describe this video briefly."}   # base_dataset.py:61-66
TVG_PROMPT = "Generate a video given the caption."                                                             # base_dataset.py:88


def chatml_prompt(messages):
    """conv_templates["qwen_2"].get_prompt() for [(role, message-or-None), ...] (conversation.py:90-100)."""
    ret = _SYSTEM + _SEP + "\n"
    for role, message in messages:
        ret += (role + "\n" + message + _SEP + "\n") if message else (role + "\n")
    return ret


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """Tokenise the text between "<image>" markers and join the pieces with the image sentinel (base_dataset.py:39-58)."""
    chunks = [tokenizer(chunk).input_ids for chunk in prompt.split(DEFAULT_IMAGE_TOKEN)]
    ids, offset = [], 0
    bos = getattr(tokenizer, "bos_token_id", None)
    if chunks and len(chunks[0]) > 0 and chunks[0][0] == bos:
        offset = 1
        ids.append(chunks[0][0])
    for i, chunk in enumerate(chunks):
        if i > 0:
            ids.append(image_token_index)   # the separator [image] * (offset + 1), sliced by `offset` like every piece
        ids.extend(chunk[offset:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return ids


class RetrievalDataset(torch.utils.data.Dataset):
    """BaseDataset (base_dataset.py:11): `data` = list of {"vid", "text"}; features are `[n_clips, 64, 1024]` tensors in
    `{root}/{dataset}/features/{vid}.pth` (zeros when the file is missing, base_dataset.py:26-31)."""

    def __init__(self, args, tokenizer=None, image_processor=None, split=None, root="./data"):
        self.args = args
        self.tokenizer = tokenizer
        self.image_processor = image_processor
        self.split = split
        self.root = root
        self.feature_dir = f"{root}/{args.dataset}/features"
        self.features = set(glob.glob(f"{self.feature_dir}/*.pth"))
        self.tvg_prefix_length = self.get_tvg_prefix_length(TVG_PROMPT)
        self.data = []

    # -- prompts
    def get_tvg_prefix_length(self, init_prompt):
        prompt = chatml_prompt([(_ROLES[0], init_prompt)])
        return len(tokenizer_image_token(prompt, self.tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt")) - 2   # base_dataset.py:20-24

    def tokenizer_image_token(self, prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
        return tokenizer_image_token(prompt, tokenizer, image_token_index, return_tensors)

    def _ids_labels(self, user_message, answer):
        tok = self.tokenizer
        prompt_ids = tokenizer_image_token(chatml_prompt([(_ROLES[0], user_message), (_ROLES[1], None)]), tok, IMAGE_TOKEN_INDEX, "pt")
        input_ids = tokenizer_image_token(chatml_prompt([(_ROLES[0], user_message), (_ROLES[1], answer)]), tok, IMAGE_TOKEN_INDEX, "pt")
        assert (prompt_ids != input_ids[:len(prompt_ids)]).sum() == 0
        labels = copy.deepcopy(input_ids)
        labels[:len(prompt_ids)] = IGNORE_INDEX
        masks = input_ids.ne(tok.pad_token_id).long()
        return input_ids, labels, masks

    def get_vtg_id(self, item):
        """base_dataset.py:60-85"""
        return self._ids_labels(f"{DEFAULT_IMAGE_TOKEN}\n{VTG_PROMPTS[self.args.dataset]}", item["text"])

    def get_tvg_id(self, item):
        """base_dataset.py:87-107"""
        return self._ids_labels(f"{TVG_PROMPT}\nCaption: {item['text']}", DEFAULT_IMAGE_TOKEN)

    # -- features
    def load_video_feature(self, vid):
        path = f"{self.feature_dir}/{vid}.pth"
        if path not in self.features:
            return torch.zeros(4, 64, 1024)
        return torch.load(path, weights_only=True)

    def get_video_vocab(self):
        """base_dataset.py:33-37 (host version, kept for the drop-in attribute; stage_corpus builds it on the device)."""
        vids = sorted(set(d["vid"] for d in self.data))
        return vids, torch.stack([self.load_video_feature(v).mean(1) for v in vids], dim=0)

    def finish(self, host_vocab=True):
        self.vids = sorted(set(d["vid"] for d in self.data))
        self._vid_index = {v: i for i, v in enumerate(self.vids)}
        self.video_vocab = self.get_video_vocab()[1] if host_vocab else None

    # -- Dataset protocol
    def __getitem__(self, idx):
        item = self.data[idx]
        vtg_ids, vtg_labels, vtg_masks = self.get_vtg_id(item)
        tvg_ids, tvg_labels, tvg_masks = self.get_tvg_id(item)
        return {"vid": item["vid"], "video": self.load_video_feature(item["vid"]), "vtg_ids": vtg_ids, "vtg_labels": vtg_labels,
                "vtg_masks": vtg_masks, "tvg_ids": tvg_ids, "tvg_labels": tvg_labels, "tvg_masks": tvg_masks,
                "tvg_video_labels": self._vid_index[item["vid"]]}

    def __len__(self):
        return len(self.data)

    def collate_fn(self, batch):
        """base_dataset.py:119-163: ragged lists at evaluation, left-padded tensors for split == 'train'."""
        out = {"vid": [b["vid"] for b in batch], "video": [b["video"] for b in batch]}
        for kind in ("vtg", "tvg"):
            ids, labels, masks = ([b[f"{kind}_{f}"] for b in batch] for f in ("ids", "labels", "masks"))
            if self.split == "train":
                n, width = len(batch), max(len(x) for x in ids)
                ids_p = torch.full((n, width), self.tokenizer.pad_token_id, dtype=torch.long)
                lab_p = torch.full((n, width), IGNORE_INDEX, dtype=torch.long)
                msk_p = torch.zeros((n, width), dtype=torch.long)
                for i in range(n):
                    m = len(ids[i])
                    ids_p[i, width - m:], lab_p[i, width - m:], msk_p[i, width - m:] = ids[i], labels[i], masks[i]
                ids, labels, masks = ids_p, lab_p, msk_p
            out[f"{kind}_ids"], out[f"{kind}_labels"], out[f"{kind}_masks"] = ids, labels, masks
        out["tvg_video_labels"] = torch.tensor([self._vid_index[b["vid"]] for b in batch])
        return out


def _keep(ds, vid):
    return ds.split == "test" or (ds.split == "train" and f"{ds.feature_dir}/{vid}.pth" in ds.features)


class MSRVTT(RetrievalDataset):
    def __init__(self, args=None, tokenizer=None, image_processor=None, split="train", root="./data", host_vocab=True):
        super().__init__(args, tokenizer, image_processor, split, root)
        self.annotations = json.load(open(f"{root}/{args.dataset}/msrvtt_ret_{split}.json"))
        for anno in self.annotations:                                   # msrvtt.py:9-13
            vid = anno["video"].split(".")[0]
            if _keep(self, vid):
                self.data.append({"vid": vid, "text": anno["caption"].strip()})
        self.finish(host_vocab)


class DiDeMo(RetrievalDataset):
    def __init__(self, args=None, tokenizer=None, image_processor=None, split="train", root="./data", host_vocab=True):
        super().__init__(args, tokenizer, image_processor, split, root)
        self.annotations = json.load(open(f"{root}/{args.dataset}/didemo_ret_{split}.json"))
        for anno in self.annotations:                                   # didemo.py:10-14: paragraph = captions joined by " "
            vid = anno["video"].split(".")[0]
            if _keep(self, vid):
                self.data.append({"vid": vid, "text": " ".join(anno["caption"]).strip()})
        self.finish(host_vocab)


class ActivityNet(RetrievalDataset):
    def __init__(self, args=None, tokenizer=None, image_processor=None, split="train", root="./data", host_vocab=True):
        super().__init__(args, tokenizer, image_processor, split, root)
        name = "anet_ret_train.json" if split == "train" else "anet_ret_val_1.json"
        self.annotations = json.load(open(f"{root}/{args.dataset}/{name}"))
        for anno in self.annotations:                                   # activitynet.py:11-15: captions joined WITHOUT a separator
            vid = anno["video"].split(".")[0]
            if _keep(self, vid):
                self.data.append({"vid": vid, "text": "".join(anno["caption"]).strip()})
        self.finish(host_vocab)


class LSMDC(RetrievalDataset):
    def __init__(self, args=None, tokenizer=None, image_processor=None, split="train", root="./data", host_vocab=True):
        super().__init__(args, tokenizer, image_processor, split, root)
        name = "lsmdc_ret_train.json" if split == "train" else "lsmdc_ret_test_1000.json"
        self.annotations = json.load(open(f"{root}/{args.dataset}/{name}"))
        for anno in self.annotations:                                   # lsmdc.py:13-16: "<movie>/<clip>.avi" -> "<clip>"
            vid = anno["video"][:-4].split("/")[1]
            if _keep(self, vid):
                self.data.append({"vid": vid, "text": anno["caption"].strip()})
        self.finish(host_vocab)


DATASETS = {"MSRVTT": MSRVTT, "DiDeMo": DiDeMo, "ActivityNet": ActivityNet, "LSMDC": LSMDC}


def load_data(args, tokenizer=None, image_processor=None, split="test", root="./data", host_vocab=True):
    """dataloader/__init__.py:8-19, evaluation branch: batches of args.batch_size_eval in dataset order."""
    dataset = DATASETS[args.dataset](args=args, tokenizer=tokenizer, image_processor=image_processor, split=split, root=root,
                                     host_vocab=host_vocab)
    return torch.utils.data.DataLoader(dataset, batch_size=args.batch_size_eval, num_workers=getattr(args, "num_workers", 0),
                                       collate_fn=dataset.collate_fn, shuffle=split == "train", pin_memory=getattr(args, "pin_mem", False),
                                       drop_last=False)


class StagedCorpus:
    """Everything evaluation() collects from the loader (retrieval_utils.py:176-197), already where the engine wants it."""

    def __init__(self, n, n_clips, tvg_video_labels, tvg_prefix_length, host_video):
        self.n, self.n_clips, self.tvg_video_labels, self.tvg_prefix_length, self.host_video = n, n_clips, tvg_video_labels, tvg_prefix_length, host_video


def stage_corpus(model, dataset, pin=True):
    """Device-side input pipeline: every item's feature file is read ONCE into one pinned host tensor `[N, n_clips, 64, MM]`
    (in the files' own 16-bit dtype -- fp16 in the reference, extract.py:107-110 -- so nothing is re-rounded), uploaded with a single asynchronous copy, the video vocabulary `feature.mean(1)` of the sorted video ids is
    reduced on the device (blim_build_video_vocab), and both token tables go to the engine as flat ragged arrays.  After
    this call score_pairs / evaluation need no per-row host work (reference: per-row list comprehension + H2D copies,
    retrieval_utils.py:55-60, and one torch.load per __getitem__ plus one per vocabulary entry, base_dataset.py:26-37)."""
    m = getattr(model, "module", model)
    eng = m.engine
    n = len(dataset)
    first = dataset.load_video_feature(dataset.data[0]["vid"])
    n_clips = first.shape[0]
    host = torch.empty((n,) + tuple(first.shape), dtype=first.dtype if first.dtype in (torch.float16, torch.bfloat16) else torch.float16)
    if pin and torch.cuda.is_available():
        host = host.pin_memory()
    vtg, tvg, labels = [], [], []
    for i, item in enumerate(dataset.data):
        host[i].copy_(first if i == 0 else dataset.load_video_feature(item["vid"]))
        vtg.append(dataset.get_vtg_id(item))
        tvg.append(dataset.get_tvg_id(item))
        labels.append(dataset._vid_index[item["vid"]])
    labels = torch.tensor(labels)
    eng.set_videos(host)
    eng.set_texts(0, [x[0] for x in vtg], [x[1] for x in vtg], [x[2] for x in vtg])
    eng.set_texts(1, [x[0] for x in tvg], [x[1] for x in tvg], [x[2] for x in tvg])
    eng.build_video_vocab(labels.numpy(), n_vocab=len(dataset.vids))
    m.set_tvg_prefix_length(dataset.tvg_prefix_length)
    m._corpus_keys.clear()   # whatever ensure_videos / ensure_texts cached is no longer what the engine holds
    dataset.staged_corpus = StagedCorpus(n, n_clips, labels, dataset.tvg_prefix_length, host)   # evaluation() skips its loader loop
    return dataset.staged_corpus
