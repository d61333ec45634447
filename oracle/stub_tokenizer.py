"""TEST INFRASTRUCTURE -- a deterministic stand-in for the Qwen2 tokenizer (no tokenizer files exist offline).

Splits on the ChatML special tokens, newlines, words and punctuation; special tokens get the real Qwen2 ids
(<|im_start|> 151644, <|im_end|> 151645, "\\n" 198, pad <|endoftext|> 151643), everything else a stable CRC32-derived id.
Used identically by the reference's BaseDataset and by blim_b200.dataset in the parity test / golden generator, so the
comparison exercises the prompt templates, the image-sentinel splicing and the label / mask rules, not the tokenizer.
"""
import re
import zlib
from types import SimpleNamespace

SPECIAL = {"<|im_start|>": 151644, "<|im_end|>": 151645, "\n": 198}
_PIECES = re.compile(r"<\|im_start\|>|<\|im_end\|>|\n|[A-Za-z0-9']+|[^\sA-Za-z0-9']")


class StubTokenizer:
    pad_token_id = 151643
    bos_token_id = None

    def __init__(self, bos=None, special=None, lo=1000, span=140000, pad=None):
        """special / lo / span / pad: remap into a small vocabulary (tiny test models); defaults = Qwen2 ids."""
        self.bos_token_id = bos
        self.special = dict(SPECIAL if special is None else special)
        self.lo, self.span = lo, span
        if pad is not None:
            self.pad_token_id = pad

    def __call__(self, text):
        ids = [] if self.bos_token_id is None else [self.bos_token_id]
        for piece in _PIECES.findall(text):
            ids.append(self.special[piece] if piece in self.special else self.lo + zlib.crc32(piece.encode()) % self.span)
        return SimpleNamespace(input_ids=ids)
