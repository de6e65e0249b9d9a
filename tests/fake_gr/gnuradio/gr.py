"""Minimal stand-in for `gnuradio.gr`: records what a block registers, publishes, tags and consumes so a
test can play the scheduler."""
sizeof_gr_complex = 8


class io_signature:
    def __init__(self, mn, mx, size):
        self.min, self.max, self.size = mn, mx, size


class basic_block:
    def __init__(self, name, in_sig, out_sig):
        self.name, self.in_sig, self.out_sig = name, in_sig, out_sig
        self.msg_in, self.msg_out, self.handlers = [], [], {}
        self.published, self.tags, self.consumed = {}, [], {}
        self.written = {}

    def message_port_register_in(self, port):
        self.msg_in.append(str(port))

    def message_port_register_out(self, port):
        self.msg_out.append(str(port))
        self.published[str(port)] = []

    def set_msg_handler(self, port, fn):
        self.handlers[str(port)] = fn

    def message_port_pub(self, port, msg):
        self.published[str(port)].append(msg)

    def add_item_tag(self, port, offset, key, value):
        self.tags.append((port, int(offset), str(key), value))

    def nitems_written(self, port):
        return self.written.get(port, 0)

    def consume(self, port, n):
        self.consumed[port] = self.consumed.get(port, 0) + int(n)


class hier_block2:
    def __init__(self, name, in_sig, out_sig):
        self.name, self.in_sig, self.out_sig = name, in_sig, out_sig
        self.connections, self.msg_connections, self.hier_in, self.hier_out = [], [], [], []

    def connect(self, a, b):
        self.connections.append((a, b))

    def msg_connect(self, a, b):
        self.msg_connections.append((a, b))

    def message_port_register_hier_in(self, port):
        self.hier_in.append(str(port))

    def message_port_register_hier_out(self, port):
        self.hier_out.append(str(port))
