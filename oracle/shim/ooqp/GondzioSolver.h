// oracle/shim: OOQP is included but never used by the reference's DDP path (ddp_optimizer.h:4-8). Empty on purpose.
