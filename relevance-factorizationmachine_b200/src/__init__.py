"""Import-path compatibility with the reference: ``from src.fm import FactorizationMachines``."""
