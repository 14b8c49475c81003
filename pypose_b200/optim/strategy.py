"""Damping strategies (reference: pypose/optim/strategy.py).

`update(pg, last, loss, J, D, R)` keeps the reference signature.  The structured LM paths never
materialise J; they pass `predicted=(J D)^T (2R + J D)` (reduced on the device) instead, which
is the only way J, D and R enter the step-quality ratio (strategy.py:143, 260).
"""


def _quality(last, loss, J, D, R, predicted):
    if predicted is None:
        JD = J @ D
        predicted = (JD.mT @ (2 * R + JD)).squeeze()
    return (last - loss) / -predicted


class Constant(object):
    """Constant damping (strategy.py:5-46)."""

    def __init__(self, damping=1e-6):
        assert damping > 0, ValueError("damping has to be positive: {}".format(damping))
        self.defaults = {'damping': damping}
    needs_quality = False

    def update(self, pg, *args, **kwargs):
        pg['damping'] = pg['damping']


class Adaptive(object):
    """Scale damping by `down` / 1 / `up` on step quality (strategy.py:49-151)."""
    needs_quality = True

    def __init__(self, damping=1e-6, high=0.5, low=1e-3, up=2., down=.5, min=1e-6, max=1e16):
        assert damping > 0, ValueError("damping has to be positive: {}".format(damping))
        assert high > 0, ValueError("high has to be positive: {}".format(high))
        assert low > 0, ValueError("low for decrease has to be positive: {}".format(low))
        assert 0 < down < 1, ValueError("down factor has to be smaller than 1: {}".format(down))
        assert 1 < up, ValueError("up factor has to be larger than 1: {}".format(up))
        self.defaults = {'damping': damping, 'high': high, 'low': low, 'up': up, 'down': down}
        self.min, self.max = min, max

    def update(self, pg, last, loss, J=None, D=None, R=None, *args, predicted=None, **kwargs):
        quality = _quality(last, loss, J, D, R, predicted)
        if quality > pg['high']:
            pg['damping'] = pg['damping'] * pg['down']
        elif quality > pg['low']:
            pg['damping'] = pg['damping']
        else:
            pg['damping'] = pg['damping'] * pg['up']
        pg['damping'] = max(self.min, min(pg['damping'], self.max))


class TrustRegion(object):
    """Trust-region radius control, damping = 1/radius (strategy.py:154-274)."""
    needs_quality = True

    def __init__(self, radius=1e6, high=.5, low=1e-3, up=2., down=.5, factor=.5, min=1e-6, max=1e16):
        assert radius > 0, ValueError("trust region radius has to be positive: {}".format(radius))
        assert high > 0, ValueError("high has to be positive: {}".format(high))
        assert low > 0, ValueError("low for decrease has to be positive: {}".format(low))
        assert 0 < down < 1, ValueError("down factor has to be smaller than 1: {}".format(down))
        assert 1 < up, ValueError("up factor has to be larger than 1: {}".format(up))
        assert 0 < factor < 1, ValueError("factor has to be smaller than 1: {}".format(factor))
        self.min, self.max, self.down = min, max, down
        self.defaults = {'radius': radius, 'damping': 1 / radius, 'high': high, 'low': low, 'up': up,
                         'down': down, 'factor': factor}

    def update(self, pg, last, loss, J=None, D=None, R=None, *args, predicted=None, **kwargs):
        quality = _quality(last, loss, J, D, R, predicted)
        pg['radius'] = 1. / pg['damping']
        if quality > pg['high']:
            pg['radius'], pg['down'] = pg['up'] * pg['radius'], self.down
        elif quality > pg['low']:
            pg['down'] = self.down
        else:
            pg['radius'] = pg['radius'] * pg['down']
            pg['down'] = pg['down'] * pg['factor']
        pg['down'] = max(self.min, min(pg['down'], self.max))
        pg['radius'] = max(self.min, min(pg['radius'], self.max))
        pg['damping'] = 1. / pg['radius']
