"""Minimal legacy-torchtext (< 0.9) stand-in: ``torchtext.data.Field`` and ``torchtext.vocab.Vocab``.

TEST INFRASTRUCTURE (see oracle/pyg_shim/README.md for the pattern): lets the UNMODIFIED reference loader
``gqa_dataset_entry.py`` be imported in the build container, where torchtext, spaCy and the GloVe download are
absent.  Only what that file touches at import time and in ``GQA_gt_sg_feature_lookup`` is provided:

* ``Field(...)`` records its special tokens (defaults ``<unk>`` / ``<pad>``); the tokenizer is never built (the
  scene-graph side passes ready-made token lists);
* ``Field.build_vocab(*sources, vectors=...)`` counts tokens the way legacy torchtext does (every element of every
  source is one example = one token list) and builds ``Vocab(counter, specials=[unk, pad, init, eos])``;
  ``vectors`` (a GloVe name) is ignored -- the embedding rows are not part of the collate path;
* ``Vocab``: ``itos`` = specials first, then tokens by descending frequency with ties in alphabetical order;
  ``stoi`` maps unknown tokens to the index of ``<unk>``.

The ordering rule is the published behaviour of torchtext 0.4-0.8 (``Vocab.__init__``: sort by token, then stable
sort by frequency, descending); it is restated here, not copied.
"""
from . import data, vocab  # noqa: F401
