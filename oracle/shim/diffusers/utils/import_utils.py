def is_xformers_available():
    # must be True: the reference defines SPLIT_SIZE only under this branch (txt_con_fusion.py:7-13)
    return True
