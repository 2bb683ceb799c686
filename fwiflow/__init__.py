"""fwiflow.jl_b200 -- B200-native FWI hot path of lidongzh/FwiFlow.jl."""
