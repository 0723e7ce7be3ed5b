"""Hotword list preparation for the contextual / SeACo Paraformer plugins — host string work, as in the reference.

Restates `generate_hotwords_list` (funasr/models/contextual_paraformer/model.py:528-660; the SeACo model carries the same
function, seaco_paraformer/model.py:583-690): a hotword source is a local `.txt` file (one hotword per line) or a string of
space-separated hotwords; when the model directory (the directory of the frontend's `cmvn_file`) holds a `seg_dict`, every word
goes through `seg_tokenize` (lower-cased dictionary lookup, CJK/digit words fall back to per-character lookup, anything else to
`<unk>`) before `tokenizer.tokens2ids`; the list always ends with the `[sos]` "no bias" entry.  URL sources need network access
and are rejected.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional

_CJK_OR_DIGITS = re.compile(r"^[一-龥0-9]+$")


def load_seg_dict(seg_dict_file: str) -> Dict[str, str]:
    """contextual_paraformer/model.py:541-557: `key piece piece ...` per line."""
    seg_dict: Dict[str, str] = {}
    with open(seg_dict_file, "r", encoding="utf8") as f:
        for line in f.readlines():
            s = line.strip().split()
            if not s:
                continue
            seg_dict[s[0]] = " ".join(s[1:])
    return seg_dict


def seg_tokenize(txt: List[str], seg_dict: Dict[str, str]) -> List[str]:
    """contextual_paraformer/model.py:559-581."""
    out_txt = ""
    for word in txt:
        word = word.lower()
        if word in seg_dict:
            out_txt += seg_dict[word] + " "
        elif _CJK_OR_DIGITS.match(word):
            for char in word:
                out_txt += (seg_dict[char] if char in seg_dict else "<unk>") + " "
        else:
            out_txt += "<unk>" + " "
    return out_txt.strip().split()


def generate_hotwords_list(hotword_list_or_file: Optional[str], tokenizer, frontend, sos: int) -> Optional[List[List[int]]]:
    """-> list of token-id lists ending with [sos], or None (no hotwords given)."""
    seg_dict = None
    cmvn_file = getattr(frontend, "cmvn_file", None)
    if cmvn_file is not None:                                            # :583-590
        seg_dict_file = os.path.join(os.path.dirname(cmvn_file), "seg_dict")
        if os.path.exists(seg_dict_file):
            seg_dict = load_seg_dict(seg_dict_file)
    if hotword_list_or_file is None:
        return None

    def ids_of(words: List[str]) -> List[int]:
        if seg_dict is not None:
            words = seg_tokenize(words, seg_dict)
        return tokenizer.tokens2ids(words)

    if os.path.exists(hotword_list_or_file) and hotword_list_or_file.endswith(".txt"):   # :594-613
        hotword_list = []
        with open(hotword_list_or_file, "r", encoding="utf8") as fin:
            for line in fin.readlines():
                hotword_list.append(ids_of(line.strip().split()))
        hotword_list.append([sos])
        return hotword_list
    if hotword_list_or_file.startswith("http"):                                          # :615-640 downloads the list
        raise ValueError("hotword lists from a URL need network access; pass a local .txt file or a string")
    if not hotword_list_or_file.endswith(".txt"):                                        # :642-653
        hotword_list = []
        for hw in hotword_list_or_file.strip().split():
            hotword_list.append(ids_of(hw.strip().split()))
        hotword_list.append([sos])
        return hotword_list
    return None
