from rfm_b200.fm import FactorizationMachines  # noqa: F401
