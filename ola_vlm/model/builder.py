"""Drop-in path of ola_vlm/model/builder.py:26-191 for the checkpoints this package's trainer writes (and the
reference's own full-model checkpoints, whose config keys and tensor names are the same): the loader that
eval / demo code calls.  Returns the reference's 4-tuple.  Generation itself is out of scope;
the loaded model serves forward passes — loss, logits, hidden states, head embeddings."""
from visper_lm_b200.model.loader import load_pretrained_model  # noqa: F401
