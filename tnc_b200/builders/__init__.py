from .circuit_builder import Circuit, Permutor
from .random_circuit import (random_circuit, random_circuit_builder, random_circuit_with_observable,
                             random_circuit_with_set_observable, random_sparse_tensor_data_with_rng)
from .sycamore_circuit import sycamore_circuit

__all__ = ["Circuit", "Permutor", "random_circuit", "random_circuit_builder", "random_circuit_with_observable",
           "random_circuit_with_set_observable", "random_sparse_tensor_data_with_rng", "sycamore_circuit"]
