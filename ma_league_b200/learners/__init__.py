from .q_learner import QLearner

# reference registry: marl/learners/__init__.py:5-9 ("sfs", "coma" are other algorithms, out of scope)
REGISTRY = {"q": QLearner}
