"""``python3 run.py --config <json>`` -- same CLI, same JSON schema and the same substring dispatch on
``config['model']['name']`` as ``/root/reference/run.py`` (:37-87).  Base models / extensions outside
the B200 hot path raise NotImplementedError instead of silently degrading.

Optional extra section the reference ignores:
  "b200": {"precision": "fp32"|"tf32"|"tf32x3", "device": "cuda:0", "schedule_seed": 123,
           "init_seed": 123, "cuda_graphs": true, "verbose": true}
  "dataset": {..., "synthetic": {"shape": "Taobao-10", "scale": 1.0, "signal": 1.0}}
"""
import argparse
import json


def in_name_list(x, name_list):
    for n in name_list:
        if n in x:
            return True
    return False


def build(config, dataset=None):
    """Dataset + base model + wrapper stack of run.py:32-65."""
    from mamdr_b200.dataset import MultiDomainDataset
    from mamdr_b200.deepctr import DeepCTR
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR

    name = config['model']['name']
    if dataset is None:
        dataset = MultiDomainDataset(config['dataset'], device=config.get('b200', {}).get('device', 'cuda:0'))

    deep_ctr_list = ['mlp', 'wdl', 'nfm', 'autoint', 'ccpm', 'pnn', 'deepfm']
    mtl_deep_ctr_list = ['shared_bottom', 'mmoe', 'ple']
    if 'star' in name:
        from mamdr_b200.star import Star
        model = Star(dataset, config)
    elif in_name_list(name, deep_ctr_list):
        model = DeepCTR(dataset, config)
    elif in_name_list(name, mtl_deep_ctr_list):
        from mamdr_b200.deep_mtl_ctr import DeepMTLCTR
        model = DeepMTLCTR(dataset, config)
    else:
        print("model: {} not found".format(name))
        raise ValueError("model: {} not found".format(name))

    if "uncertainty_weight" in name:
        raise NotImplementedError("uncertainty_weight is out of scope (baseline method, SURVEY.md 2.1 #10)")
    if "pcgrad" in name:
        from mamdr_b200.pcgrad import PCGrad
        model = PCGrad(model)

    if "meta" in name:
        if "domain_negotiation" in name:
            model = DomainNegotiation(model)
        elif "mamdr" in name:
            model = MAMDR(model)
        elif "reptile" in name:
            from mamdr_b200.reptile import Reptile
            model = Reptile(model)
        elif "mldg" in name:
            from mamdr_b200.mldg import MLDG
            model = MLDG(model)
        else:
            from mamdr_b200.maml import MAML
            model = MAML(model)
    return model


def main(config):
    # torchrun: one process per GPU.  The process group is initialised from the environment, every rank takes the GPU of its
    # LOCAL_RANK, and only rank 0 writes checkpoints / results (mamdr_b200/base_model.py: save_model, save_result).
    import os
    from mamdr_b200 import dist as mdist
    rank, world = mdist.init_from_env()
    if world > 1:
        import torch
        local = int(os.environ.get("LOCAL_RANK", rank))
        config.setdefault('b200', {})['device'] = "cuda:%d" % local
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
    model = build(config)
    name = config['model']['name']

    # Train Model
    if "separate" in name:
        avg_loss, avg_auc, domain_loss, domain_auc = model.separate_train_val_test()
    else:
        model.train()
        print("Test Result: ")
        avg_loss, avg_auc, domain_loss, domain_auc = model.val_and_test("test")

    # Finetune the model on different domains
    if "finetune" in name:
        model.load_model(model.checkpoint_path)
        print("Finetune: ")
        avg_loss, avg_auc, domain_loss, domain_auc = model.separate_train_val_test(init_parms=False)

    model.save_result(avg_loss, avg_auc, domain_loss, domain_auc)
    return avg_loss, avg_auc, domain_loss, domain_auc


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, help="Train config file", required=True)
    args = parser.parse_args()
    with open(args.config, 'r') as f:
        config = json.load(f)
    main(config)
