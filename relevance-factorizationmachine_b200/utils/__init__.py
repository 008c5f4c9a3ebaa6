"""Import-path compatibility with the reference: ``from utils.evaluate import TestEvaluator``."""
