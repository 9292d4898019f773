"""A deterministic 2048-token stand-in tokenizer whose re-tokenisation does NOT reproduce generated tokens: the test bed of
the BPE-repair heuristic (code_base/arithmetic.py:300-342) and of text_to_spans.

ids 0..25      'a'..'z'
ids 26..701    two-letter tokens 'aa'..'zz'  (id 700 is "." and id 701 is "!" instead: sentence ends)
ids 702..2044  four-digit tokens "0702".."2044" (their own id, so digits parse back uniquely); from 1500 on the text
               carries a trailing "." -- a quarter of the vocabulary ends a sentence, so finish_sent tails are short
ids 2045, 2046 "<" and ">"
id  2047       "<|endoftext|>" (the coder forbids the last id)
``encode`` is greedy longest match, except that it never merges a pair starting with 'q' and never emits the coder's
forbidden id 628 -- so a generated two-letter token may come back as two letters, and two generated letters as a pair.
"""
import string

V = 2048
LETTERS = string.ascii_lowercase


class StubTokenizer:
    def __init__(self):
        self.text = {}
        for i, ch in enumerate(LETTERS):
            self.text[i] = ch
        k = 26
        for a in LETTERS:
            for b in LETTERS:
                self.text[k] = a + b
                k += 1
        self.text[700], self.text[701] = ".", "!"
        for i in range(702, 2045):
            self.text[i] = ("%04d." if i >= 1500 else "%04d") % i
        self.text[2045] = "<"
        self.text[2046] = ">"
        self.text[2047] = "<|endoftext|>"
        self.pair = {t: i for i, t in self.text.items() if len(t) == 2 and t.isalpha()}
        self.single = {t: i for i, t in self.text.items() if len(t) == 1}      # letters, ".", "!", "<", ">"
        self.vocab_size = V

    def decode(self, ids, skip_special_tokens=True):
        return "".join("" if (skip_special_tokens and int(i) == 2047) else self.text[int(i)] for i in ids)

    def encode(self, text, add_special_tokens=False):
        if text == "<|endoftext|>":
            return [2047]
        out, p = [], 0
        while p < len(text):
            ch = text[p]
            if ch.isdigit():
                n = int(text[p:p + 4])
                out.append(n); p += 5 if n >= 1500 else 4
                continue
            two = text[p:p + 2]
            if len(two) == 2 and two in self.pair and ch != "q" and self.pair[two] != 628:
                out.append(self.pair[two]); p += 2
                continue
            out.append(self.single[ch]); p += 1
        return out
