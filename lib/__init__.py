"""Import-path shim: zju3dv/pvnet's entry points import `lib.ransac_voting_gpu_layer.*`
and `lib.networks.model_repository`; these packages re-export pvnet_b200's drop-ins
under those names."""
