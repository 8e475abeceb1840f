"""Host-side caches of the product path (no GPU needed): the parameter lists / tables handed to the C ABI are cached
per module and must follow parameter replacement, in-place updates and storage moves."""
import torch
import torch.nn as nn


def test_smt_parameter_references_follow_replacement_and_loading():
    from avlen_b200.savi.models.smt_state_encoder import SMT_PARAM_KEYS, SMTStateEncoder
    enc = SMTStateEncoder(276, dim_feedforward=256, pose_indices=(272, 276), nhead=8, num_encoder_layers=1,
                          num_decoder_layers=1, dropout=0.0, activation="relu")
    named = dict(enc.named_parameters())
    first = enc._params()
    assert len(first) == len(SMT_PARAM_KEYS)
    for k, p in zip(SMT_PARAM_KEYS, first):
        assert p is named[k]
    # a parameter object replaced inside its module is picked up (the cache holds (owner dict, name), not tensors)
    new_w = nn.Parameter(torch.zeros_like(enc.pose_encoder.weight))
    enc.pose_encoder.weight = new_w
    again = enc._params()
    assert again[SMT_PARAM_KEYS.index("pose_encoder.weight")] is new_w
    # load_state_dict copies in place: same objects, bumped versions
    v0 = new_w._version
    enc.load_state_dict({k: torch.ones_like(v) for k, v in enc.state_dict().items()})
    assert enc._params()[SMT_PARAM_KEYS.index("pose_encoder.weight")] is new_w and new_w._version > v0


def test_resnet_plan_fingerprint_sees_updates_and_storage_moves():
    from avlen_b200.savi.models.smt_resnet import custom_resnet18
    net = custom_resnet18(num_input_channels=3)
    plan = net.plan()
    fp0 = plan._fingerprint(1)
    assert plan._fingerprint(1) == fp0 and plan._fingerprint(0) != fp0
    with torch.no_grad():
        net.layer2[0].conv1.weight.mul_(0.5)          # in-place update (optimizer step, load_state_dict)
    fp1 = plan._fingerprint(1)
    assert fp1 != fp0
    net.fc.bias.data = net.fc.bias.data.clone()       # storage move (.to(), flatten_parameters)
    fp2 = plan._fingerprint(1)
    assert fp2 != fp1
    net.conv1.weight = nn.Parameter(net.conv1.weight.detach().clone())   # parameter object replaced
    assert plan._fingerprint(1) != fp2
