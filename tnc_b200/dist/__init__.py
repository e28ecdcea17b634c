from .communication import (Communication, PartitionedPlan, RankTensorMapping, broadcast_path, broadcast_serializing, contract_partitioned,
                            fanin_schedule, get_tensor_mapping, init_device_comm,
                            intermediate_reduce_tensor_network, scatter_tensor_network)

__all__ = ["Communication", "PartitionedPlan", "RankTensorMapping", "broadcast_path", "broadcast_serializing", "contract_partitioned", "fanin_schedule",
           "get_tensor_mapping", "init_device_comm", "intermediate_reduce_tensor_network", "scatter_tensor_network"]
