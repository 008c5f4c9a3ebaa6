from rfm_b200.base import PointwiseBaseRecommender  # noqa: F401
