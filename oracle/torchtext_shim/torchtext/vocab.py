"""``torchtext.vocab.Vocab`` as legacy torchtext builds it (restated; see the package docstring)."""
import collections


class Vocab:
    UNK = "<unk>"

    def __init__(self, counter, max_size=None, min_freq=1, specials=("<unk>", "<pad>"), vectors=None, **_unused):
        self.freqs = counter
        counts = collections.Counter(counter)
        specials = list(specials)
        for tok in specials:                      # specials never compete for a frequency rank
            counts.pop(tok, None)
        ranked = sorted(counts.items(), key=lambda kv: kv[0])          # alphabetical ...
        ranked.sort(key=lambda kv: kv[1], reverse=True)                # ... then stable by descending frequency
        self.itos = list(specials)
        limit = None if max_size is None else max_size + len(specials)
        for tok, freq in ranked:
            if freq < max(min_freq, 1) or (limit is not None and len(self.itos) >= limit):
                break
            self.itos.append(tok)
        unk = specials.index(self.UNK) if self.UNK in specials else None
        self.unk_index = unk
        self.stoi = collections.defaultdict((lambda: unk) if unk is not None else None)
        self.stoi.update({tok: i for i, tok in enumerate(self.itos)})
        self.vectors = None                       # `vectors=` (GloVe) is not loaded by the shim

    def __len__(self):
        return len(self.itos)

    def __getitem__(self, token):
        return self.stoi.get(token, self.unk_index)
