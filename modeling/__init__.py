"""Drop-in package: ``from modeling.t2i_pipeline import BitDanceT2IPipeline`` (as in the reference's example_t2i.py,
app.py and eval/*.py) resolves to the B200-native implementation when this repository root is on ``sys.path``."""
