class TTNS:
    pass
