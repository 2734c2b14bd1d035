class Layer(object):
    """keras.layers.Layer: only what SMPLLayer touches (constructor kwargs, get_config)."""

    def __init__(self, **kwargs):
        self.name = kwargs.get("name")
        self.built = False

    def get_config(self):
        return {"name": self.name}

    def __call__(self, x):
        if not self.built:
            self.build(None)
            self.built = True
        return self.call(x)
