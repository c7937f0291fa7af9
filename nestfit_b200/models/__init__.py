from . import ammonia, gaussian

MODEL_MODULES = [ammonia, gaussian]
MODELS = {m.NAME: m for m in MODEL_MODULES}
