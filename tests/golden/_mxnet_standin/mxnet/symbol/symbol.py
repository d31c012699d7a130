class Symbol(object):
    pass
