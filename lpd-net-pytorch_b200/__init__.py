"""lpdnet_b200 — B200-native (sm_100a) drop-in for the LPD-Net / PointNetVLAD hot path.

Public surface mirrors the reference's module API (qiaozhijian/LPD-Net-Pytorch):
    lpdnet_b200.util.PointNetVlad   PointNetVlad, NetVLADLoupe, GatingContext, STN3d, PointNetfeat
    lpdnet_b200.util.lpdnet_model   LPDNet, LPDNetOrign, TranformNet, knn, get_graph_feature[_Origin]
    lpdnet_b200.loss.pointnetvlad_loss   best_pos_distance, triplet_loss[_wrapper], quadruplet_loss
    lpdnet_b200.evaluate            get_recall, get_latent_vectors
All arithmetic runs in liblpd_b200.so (include/lpd_b200.h); there is no CPU / torch fallback.

The directory is named `lpd-net-pytorch_b200` (not importable as-is); the sibling package
`lpdnet_b200` aliases it.
"""
__version__ = "0.2.0"
