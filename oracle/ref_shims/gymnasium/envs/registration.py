registry = {}


def register(id, entry_point=None, max_episode_steps=None, **kw):
    registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps)


def make(*a, **k):
    raise NotImplementedError("stand-in gymnasium: make() is not provided")


def spec(*a, **k):
    raise NotImplementedError("stand-in gymnasium: spec() is not provided")
