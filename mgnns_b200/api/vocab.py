"""Vocabulary file access (ref: utils/vocab_new.py:8-34; utils/vocab.py is identical except that it
reads `vocab_new/` — a directory the reference does not ship, so both names resolve here)."""
import os


def get_vocab(vocab_root_path, text_min_count):
    """Read `<root>/vocab/vocab-<N>.txt`, one word per line; PAD is index 0 and UNK index 1.

    Like the reference, the file is split on '\\n' without stripping, so a trailing newline yields a
    final empty-string entry that counts as a vocabulary word.
    """
    for sub in ('vocab', 'vocab_new'):
        path = os.path.join(vocab_root_path, sub, 'vocab-' + str(text_min_count) + '.txt')
        if os.path.exists(path):
            with open(path) as f:
                return f.read().split('\n')
    raise FileNotFoundError(os.path.join(vocab_root_path, 'vocab', 'vocab-' + str(text_min_count) + '.txt'))


def get_vocab_list(data_root_path, vocab_root_path, text_min_count):
    """(ref: utils/vocab_new.py:8-14).  Building a vocabulary from the train split is one-off data
    preparation outside the hot path; only reading an existing file is supported here."""
    return get_vocab(vocab_root_path, text_min_count)
