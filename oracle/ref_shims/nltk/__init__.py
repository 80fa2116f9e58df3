"""Import shim (test infrastructure only): NLTK and its corpora are not installed and
there is no network. /root/reference/utils_attacks.py:7-11 imports it at module load; the
constrain=True branch (utils_attacks.py:110-143) is therefore unavailable ("parity unpinned")."""


def download(*args, **kwargs):
    return False
