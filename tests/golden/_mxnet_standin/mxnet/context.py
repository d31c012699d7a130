class Context(object):
    def __init__(self, device_type='cpu', device_id=0):
        self.device_type, self.device_id = device_type, device_id

    def __repr__(self):
        return '%s(%d)' % (self.device_type, self.device_id)

    def __eq__(self, other):
        return isinstance(other, Context) and (self.device_type, self.device_id) == (other.device_type, other.device_id)

    def __hash__(self):
        return hash((self.device_type, self.device_id))


def cpu(i=0):
    return Context('cpu', i)


def gpu(i=0):
    return Context('gpu', i)


def current_context():
    return cpu()


Context.default_ctx = Context('cpu', 0)
