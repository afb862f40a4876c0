          seed = 12345

       seqfile = frogs.txt
      Imapfile = frogs.Imap.txt
       jobname = out

  speciesdelimitation = 0
         speciestree = 0

   species&tree = 4  K  C  L  H
                     9  7 14  2
                  (((K, C), L), H);

         phase = 0 0 0 0
       usedata = 1
         model = gtr
    alphaprior = 1 1 4
       scaling = 1
         nloci = 5
     cleandata = 0

    thetaprior = gamma 2 2000
      tauprior = gamma 2 1000

      finetune = 1

         print = 1 0 0 0
        burnin = 200
      sampfreq = 2
       nsample = 500
