class CLIPScore:  # imported, never used, by /root/reference/utils_attacks.py:6
    pass
