"""Column-expression placeholders; the local DataFrame interprets them by name."""


class Column(object):
    def __init__(self, op, arg=None):
        self.op, self.arg = op, arg


def rand(seed=None):
    return Column("rand", seed)


def mean(col):
    return Column("mean", col)


def stddev_pop(col):
    return Column("stddev_pop", col)


def col(name):
    return Column("col", name)
