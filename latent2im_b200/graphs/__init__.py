"""Graph factory (reference graphs/__init__.py:3-22)."""
import importlib


def find_model_using_name(model, transform):
    lib = importlib.import_module(__name__ + ".transform_graph_scene")
    wanted = (transform.replace("_", "") + "graph").lower()
    for g in lib.get_transform_graphs(model):
        if g.__name__.lower() == wanted:
            return g
    raise ValueError("no transform graph named %r for model %r" % (wanted, model))
