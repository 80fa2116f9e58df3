def word_tokenize(*args, **kwargs):
    raise RuntimeError("nltk shim: word_tokenize unavailable (no NLTK data offline)")
