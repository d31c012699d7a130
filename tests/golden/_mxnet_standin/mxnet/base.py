class MXNetError(Exception):
    pass
