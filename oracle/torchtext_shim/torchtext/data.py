"""``torchtext.data.Field`` subset (restated; see the package docstring)."""
import collections
import itertools

from .vocab import Vocab


class Field:
    vocab_cls = Vocab

    def __init__(self, sequential=True, use_vocab=True, init_token=None, eos_token=None, fix_length=None,
                 dtype=None, preprocessing=None, postprocessing=None, lower=False, tokenize=None,
                 tokenizer_language="en", include_lengths=False, batch_first=False, pad_token="<pad>",
                 unk_token="<unk>", pad_first=False, truncate_first=False, stop_words=None, is_target=False):
        self.sequential, self.use_vocab, self.lower = sequential, use_vocab, lower
        self.init_token, self.eos_token = init_token, eos_token
        self.pad_token = pad_token if sequential else None
        self.unk_token = unk_token
        self.include_lengths, self.batch_first = include_lengths, batch_first
        self.tokenize, self.tokenizer_language = tokenize, tokenizer_language    # recorded, never instantiated

    def build_vocab(self, *sources, **kwargs):
        counter = collections.Counter()
        for data in sources:
            for example in data:
                if not self.sequential:
                    example = [example]
                try:
                    counter.update(example)
                except TypeError:
                    counter.update(itertools.chain.from_iterable(example))
        extra = kwargs.pop("specials", [])
        specials = list(collections.OrderedDict.fromkeys(
            tok for tok in [self.unk_token, self.pad_token, self.init_token, self.eos_token] + list(extra)
            if tok is not None))
        kwargs.pop("vectors", None)
        self.vocab = self.vocab_cls(counter, specials=specials, **kwargs)
