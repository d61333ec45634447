"""Host flow of the reference's extract.py on the B200 extractor.

    chunk_bounds(n, num_chunk, chunk_idx)        extract.py:79-85   (which slice of the sorted video list a worker owns)
    extract_dataset(encoder, video_list, load_frames, out_dir, ...)   extract.py:96-109
        for every video: frames [T, 3, S, S] -> encoder.extract_videos -> torch.save(fp16 [T/4, 64, 1024], "{vid}.pth")

Decoding and preprocessing stay with the caller: `load_frames(path) -> [T, 3, S, S]` pixel tensor is what the reference's
VideoDataset.__getitem__ returns (decord + UMTImageProcessor, extract.py:42-76; neither decord nor PIL resizing is part of
the engine).  Videos whose loader raises are skipped like in the reference (extract.py:71-75 moves on to the next index).
"""
import os

import torch


def chunk_bounds(n, num_chunk, chunk_idx):
    """extract.py:79-85: equal chunks, the last one takes the remainder."""
    size = n // num_chunk
    start = size * chunk_idx
    end = n if chunk_idx == num_chunk - 1 else min(size * (chunk_idx + 1), n)
    return start, end


def video_id(path, dataset):
    """extract.py:66-69"""
    base = os.path.basename(path)
    return base[:-4] if dataset == "LSMDC" else base.split(".")[0]


def extract_dataset(encoder, video_list, load_frames, out_dir, dataset="MSRVTT", num_chunk=1, chunk_idx=0, batch_videos=4, log=None):
    """Extracts and saves the features of this worker's chunk of `video_list` (sorted like extract.py:77).  Returns the list
    of video ids written."""
    video_list = sorted(video_list)
    start, end = chunk_bounds(len(video_list), num_chunk, chunk_idx)
    os.makedirs(out_dir, exist_ok=True)
    done, pending = [], []

    def flush():
        if not pending:
            return
        feats = encoder.extract_videos([f for _, f in pending], out_dtype=torch.float16)
        for (vid, _), feat in zip(pending, feats):
            torch.save(feat.cpu(), os.path.join(out_dir, f"{vid}.pth"))     # extract.py:104-106: one fp16 tensor per video
            done.append(vid)
        pending.clear()

    fpc = encoder.cfg.frames_per_clip
    for path in video_list[start:end]:
        try:
            frames = load_frames(path)
        except Exception as ex:   # extract.py:73-75
            if log:
                log(f"Error loading video {path}: {ex}")
            continue
        clips = sum(f.shape[0] // fpc for _, f in pending) + frames.shape[0] // fpc
        if pending and (len(pending) >= batch_videos or clips > encoder.max_clips):
            flush()
        pending.append((video_id(path, dataset), frames))
    flush()
    return done
