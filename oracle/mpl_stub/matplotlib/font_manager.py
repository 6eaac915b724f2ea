"""TEST INFRASTRUCTURE -- see matplotlib/__init__.py in this directory."""


class FontProperties:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
