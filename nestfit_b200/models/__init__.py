from . import ammonia, diazenylium, gaussian

MODEL_MODULES = [ammonia, gaussian, diazenylium]
MODELS = {m.NAME: m for m in MODEL_MODULES}
