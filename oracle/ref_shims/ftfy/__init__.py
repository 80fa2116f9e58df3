"""Import shim (test infrastructure only): ftfy is not installed in this image.

The reference calls ftfy.fix_text in basic_clean (/root/reference/src/open_clip/tokenizer.py:67).
Identity is exact for the printable-ASCII captions this repo's parity domain covers,
except ftfy's extra upper-case `&NAME;` entity variants (documented in DESIGN.md)."""


def fix_text(text, *args, **kwargs):
    return text
