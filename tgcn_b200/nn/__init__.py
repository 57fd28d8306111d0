from .gcn import GCNCheb, TGCNCheb, TGCNCheb_H, gcn_pool, gcn_pool_4, relu_pool  # noqa: F401
