from .circuit_builder import Circuit, Permutor
from .random_circuit import random_circuit, random_circuit_builder
from .sycamore_circuit import sycamore_circuit

__all__ = ["Circuit", "Permutor", "random_circuit", "random_circuit_builder", "sycamore_circuit"]
