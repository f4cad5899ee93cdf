"""Keras layer descriptions (config holders); ``keras.models.Sequential`` turns them into torch modules."""


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else tuple(int(x) for x in v)


class Layer(object):
    class_name = "Layer"

    def __init__(self, input_shape=None, input_dim=None, name=None, **kwargs):
        if input_shape is None and input_dim is not None:
            input_shape = (int(input_dim),)
        self.input_shape = None if input_shape is None else tuple(int(s) for s in input_shape)
        self.name = name

    def get_config(self):
        return {}

    def _base_config(self):
        cfg = self.get_config()
        if self.input_shape is not None:
            cfg["batch_input_shape"] = [None] + list(self.input_shape)
        if self.name is not None:
            cfg["name"] = self.name
        return cfg

    @classmethod
    def from_config(cls, cfg):
        cfg = dict(cfg)
        bis = cfg.pop("batch_input_shape", None)
        if bis is not None:
            cfg["input_shape"] = tuple(bis[1:])
        return cls(**cfg)


class Dense(Layer):
    class_name = "Dense"

    def __init__(self, units, activation=None, use_bias=True, **kwargs):
        super(Dense, self).__init__(**kwargs)
        self.units, self.activation, self.use_bias = int(units), activation, bool(use_bias)

    def get_config(self):
        return {"units": self.units, "activation": self.activation or "linear", "use_bias": self.use_bias}


class Activation(Layer):
    class_name = "Activation"

    def __init__(self, activation, **kwargs):
        super(Activation, self).__init__(**kwargs)
        self.activation = activation

    def get_config(self):
        return {"activation": self.activation}


class Dropout(Layer):
    class_name = "Dropout"

    def __init__(self, rate, **kwargs):
        super(Dropout, self).__init__(**kwargs)
        self.rate = float(rate)

    def get_config(self):
        return {"rate": self.rate}


class Flatten(Layer):
    class_name = "Flatten"


class Reshape(Layer):
    class_name = "Reshape"

    def __init__(self, target_shape, **kwargs):
        super(Reshape, self).__init__(**kwargs)
        self.target_shape = tuple(int(s) for s in target_shape)

    def get_config(self):
        return {"target_shape": list(self.target_shape)}


class Conv2D(Layer):
    class_name = "Conv2D"

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", activation=None, use_bias=True, **kwargs):
        super(Conv2D, self).__init__(**kwargs)
        self.filters, self.kernel_size, self.strides = int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.activation, self.use_bias = padding, activation, bool(use_bias)

    def get_config(self):
        return {"filters": self.filters, "kernel_size": list(self.kernel_size), "strides": list(self.strides),
                "padding": self.padding, "activation": self.activation or "linear", "use_bias": self.use_bias}


def Convolution2D(filters, kh, kw=None, **kwargs):
    """Keras-1 spelling used by the reference's examples (``examples/mnist.py:151``)."""
    if kw is None:
        return Conv2D(filters, kh, **kwargs)
    if "border_mode" in kwargs:
        kwargs["padding"] = kwargs.pop("border_mode")
    return Conv2D(filters, (kh, kw), **kwargs)


class MaxPooling2D(Layer):
    class_name = "MaxPooling2D"

    def __init__(self, pool_size=(2, 2), strides=None, **kwargs):
        super(MaxPooling2D, self).__init__(**kwargs)
        self.pool_size = _pair(pool_size)
        self.strides = self.pool_size if strides is None else _pair(strides)

    def get_config(self):
        return {"pool_size": list(self.pool_size), "strides": list(self.strides)}


LAYERS = {c.class_name: c for c in (Dense, Activation, Dropout, Flatten, Reshape, Conv2D, MaxPooling2D)}
