class _Words:
    def words(self):
        raise RuntimeError("nltk shim: words corpus unavailable (no NLTK data offline)")


words = _Words()
