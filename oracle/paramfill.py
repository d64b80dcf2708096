"""Oracle (test infrastructure): the procedural weight filler lives in the product package
(``human_instance_segmentation_b200/synthetic.py``, it only generates numbers) so that the golden generator, the
oracle, the tests and the benchmark share ONE definition; re-exported here for the oracle-side callers."""
from human_instance_segmentation_b200.synthetic import fill_state_dict  # noqa: F401
