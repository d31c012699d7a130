class Initializer(object):
    pass


class Constant(Initializer):
    def __init__(self, value):
        self.value = value


class Uniform(Initializer):
    def __init__(self, scale=0.07):
        self.scale = scale


class Xavier(Initializer):
    def __init__(self, rnd_type='uniform', factor_type='avg', magnitude=3):
        self.magnitude = magnitude
