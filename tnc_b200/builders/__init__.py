from .circuit_builder import Circuit, Permutor

__all__ = ["Circuit", "Permutor"]
