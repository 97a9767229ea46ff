class Prior(object):
    def __init__(self, name=None, latex_label=None, unit=None, minimum=None, maximum=None, boundary=None):
        self.name, self.minimum, self.maximum = name, minimum, maximum
