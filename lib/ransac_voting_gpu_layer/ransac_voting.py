"""`lib.ransac_voting_gpu_layer.ransac_voting` (the reference's native extension module,
src/ransac_voting.cpp:102-107), served by pvnet_b200."""
from pvnet_b200.ransac_voting import (generate_hypothesis, generate_hypothesis_vanishing_point,  # noqa: F401
                                      voting_for_hypothesis, voting_for_hypothesis_vanishing_point)
